#include <chrono>
#include <cstdio>
#include <cstring>
#include "../kogarashi_b200/csrc/curve.cuh"
using namespace kgr;
template <class C> double bench_dbl(int iters) {
    XyzzPt<C> p;
    typedef typename C::Elem E;
    uint32_t *w = (uint32_t *)&p;
    for (size_t i = 0; i < sizeof(p) / 4; i++) w[i] = 0x12345u * (i + 1) & 0x0fffffffu;
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) p = xyzz_dbl(p);
    double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / iters;
    volatile uint32_t sink = w[0]; (void)sink;
    return ns;
}
template <class E> double bench_mul(int iters) {
    E a, b;
    uint32_t *wa = (uint32_t *)&a, *wb = (uint32_t *)&b;
    for (size_t i = 0; i < sizeof(a) / 4; i++) { wa[i] = 0x12345u * (i + 1) & 0x0fffffffu; wb[i] = 0x54321u * (i + 3) & 0x0fffffffu; }
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) a = fp_mul(a, b);
    double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count() / iters;
    volatile uint32_t sink = wa[0]; (void)sink;
    return ns;
}
int main() {
    printf("Fq mul %.1f ns, Fq2 mul %.1f ns\n", bench_mul<Fp<FqP>>(2000000), bench_mul<Fp2<FqP>>(1000000));
    printf("G1 xyzz_dbl %.1f ns, G2 xyzz_dbl %.1f ns\n", bench_dbl<Bn254G1>(300000), bench_dbl<Bn254G2>(100000));
}
