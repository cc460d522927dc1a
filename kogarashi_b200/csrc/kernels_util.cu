// kernels_util.cu — curve-independent kernels: prefix scan, field-op test hook, integer-pipe microbenchmarks.
#include <cuda_runtime.h>

#include "launch.cuh"
#include "modinv.cuh"
#include "scan.cuh"

namespace kgr {

template <class P> __global__ void k_field_op(int op, const Fp<P> *a, const Fp<P> *b, Fp<P> *out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<P> x = a[i], y = b ? b[i] : fp_zero<P>(), r;
    switch (op) {
        case 0: r = fp_add(x, y); break;
        case 1: r = fp_sub(x, y); break;
        case 2: r = fp_mul(x, y); break;
        case 3: r = fp_sqr(x); break;
        case 4: r = fp_neg(x); break;
        case 5: r = fp_from_mont(x); break;
        case 6: r = fp_to_mont(x); break;
        case 7: r = fp_inv(x); break;
        case 9: r = fp_inv_fast(x); break;
        default: r = fp_dbl(x); break;
    }
    out[i] = r;
}
// ---- integer-pipe microbenchmarks -------------------------------------------------------------
template <int MODE> __global__ void __launch_bounds__(256) k_ubench(uint32_t *sink, uint32_t a, uint32_t b, int iters) {
    uint32_t x[8];
    uint64_t y[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        x[j] = threadIdx.x * 7 + j;
        y[j] = x[j];
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 8; rep++) {
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 8; j++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a), "r"(b));
            } else if (MODE == 1) {
#pragma unroll
                for (int j = 0; j < 8; j++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a), "r"(b));
            } else if (MODE == 2) {
#pragma unroll
                for (int j = 0; j < 8; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[j]) : "r"(a), "r"(x[j]));
            } else if (MODE == 3) {
                uint32_t top = 0;
                chain_cmad(x, a, b, a ^ 0x55u, b ^ 0x33u, x[0] | 1u, top);
                x[1] ^= top;
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(a));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) acc ^= x[j] ^ (uint32_t)y[j] ^ (uint32_t)(y[j] >> 32);
    if (acc == 0x12345u) sink[0] = acc;
}
__global__ void __launch_bounds__(256) k_ubench_fmul(Fp<FqP> *sink, Fp<FqP> a, Fp<FqP> b, int iters) {
    a.v[0] ^= threadIdx.x;
    for (int it = 0; it < iters; it++) {
        a = fp_mul(a, b);
        b = fp_mul(b, a);
    }
    if (a.v[0] == 0x12345u && b.v[1] == 7u) sink[0] = a;
}
__global__ void __launch_bounds__(128) k_ubench_madd(XyzzPt<Bn254G1> *sink, AffinePt<Bn254G1> p, AffinePt<Bn254G1> q, int iters) {
    XyzzPt<Bn254G1> acc = xyzz_from_affine(p);
    acc.x.v[0] ^= (threadIdx.x & 1);  // not a curve point any more; the formulas do not care
    for (int it = 0; it < iters; it++) xyzz_madd(acc, q);
    if (acc.x.v[0] == 0x12345u && acc.y.v[1] == 7u) sink[0] = acc;
}
__global__ void k_clock(uint64_t *out) {
    uint64_t t0, c0, t1, c1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    c0 = clock64();
    do {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < 2000000ULL);
    c1 = clock64();
    out[0] = t1 - t0;
    out[1] = c1 - c0;
}


void LaunchUtil::field_op(cudaStream_t st, int field, int op, const void *a, const void *b, void *out, uint32_t n) {
    unsigned blocks = (n + 127) / 128;
    if (field == 0) k_field_op<FqP><<<blocks, 128, 0, st>>>(op, (const Fp<FqP> *)a, (const Fp<FqP> *)b, (Fp<FqP> *)out, n);
    else k_field_op<FrP><<<blocks, 128, 0, st>>>(op, (const Fp<FrP> *)a, (const Fp<FrP> *)b, (Fp<FrP> *)out, n);
}
void LaunchUtil::exclusive_scan(cudaStream_t st, const uint32_t *in, uint32_t *out, uint32_t len, uint32_t *tile_sums) { exclusive_scan_u32(in, out, len, tile_sums, st); }
uint32_t LaunchUtil::scan_tiles(uint32_t len) { return scan_num_tiles(len); }
int LaunchUtil::scan_launches(uint32_t len) { return scan_num_tiles(len) > 1 ? 3 : 1; }
void LaunchUtil::ubench(cudaStream_t st, int mode, int blocks, uint32_t *sink, int iters) {
    switch (mode) {
        case 0: k_ubench<0><<<blocks, 256, 0, st>>>(sink, 3, 5, iters); break;
        case 1: k_ubench<1><<<blocks, 256, 0, st>>>(sink, 3, 5, iters); break;
        case 2: k_ubench<2><<<blocks, 256, 0, st>>>(sink, 3, 5, iters); break;
        case 3: k_ubench<3><<<blocks, 256, 0, st>>>(sink, 3, 5, iters); break;
        default: k_ubench<4><<<blocks, 256, 0, st>>>(sink, 3, 5, iters); break;
    }
}
void LaunchUtil::ubench_fmul(cudaStream_t st, int blocks, void *sink, int iters) {
    Fp<FqP> x = fp_one<FqP>(), y = fp_one<FqP>();
    y.v[0] ^= 0x1234;
    k_ubench_fmul<<<blocks, 256, 0, st>>>((Fp<FqP> *)sink, x, y, iters);
}
void LaunchUtil::ubench_madd(cudaStream_t st, int blocks, void *sink, const AffinePt<Bn254G1> &p, const AffinePt<Bn254G1> &q, int iters) {
    k_ubench_madd<<<blocks, 128, 0, st>>>((XyzzPt<Bn254G1> *)sink, p, q, iters);
}
void LaunchUtil::clock_probe(cudaStream_t st, uint64_t *out2) { k_clock<<<1, 1, 0, st>>>(out2); }

}  // namespace kgr
