// oracle/zkstd_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see zkstd_oracle.hpp header).
// extern "C" surface over the CPU restatement so tests / bench.py's cpu_baseline can drive it
// through ctypes.  Buffer formats are the same as include/kgr_msm.h so a test can hand the same
// arrays to both sides.
#include "zkstd_oracle.hpp"
#include <thread>
#include <atomic>
#include <algorithm>

using namespace zko;

namespace {

template <class C> struct Msm {
    typedef Curve<C> Cv;
    typedef Field<typename C::Scalar> Fs;

    // groth16/src/msm.rs:50-73
    struct Bucket {
        uint8_t kind;  // 0 None, 1 Affine, 2 Projective
        Affine a;
        Proj p;
    };
    static inline void bucket_add_assign(Bucket &b, const Affine &other) {
        switch (b.kind) {
            case 0: b.a = other; b.kind = 1; break;
            case 1: b.p = Cv::add_affine(b.a, other); b.kind = 2; break;   // a + other  (Affine + &Affine)
            default: b.p = Cv::add_mixed(other, b.p); break;               // ext + aff -> add_mixed(aff, ext)
        }
    }
    static inline Proj bucket_add(const Bucket &b, const Proj &other) {
        switch (b.kind) {
            case 0: return other;
            case 1: return Cv::add_mixed(b.a, other);   // other + a (macros/curve.rs:182-188)
            default: return Cv::add_proj(other, b.p);   // other + a
        }
    }

    // groth16/src/msm.rs:6-48.  `threads` mirrors rayon's par_iter_mut over windows (:17-24).
    static Proj run(const Affine *bases, size_t n_bases, const Limbs *coeffs, size_t n_coeffs, int threads) {
        size_t c = window_bits(n_bases);
        size_t n_windows = 256 / c + 1;
        size_t n_pairs = std::min(n_bases, n_coeffs);  // zip semantics (:25)
        std::vector<Proj> window_acc(n_windows);
        std::atomic<size_t> next(0);
        auto worker = [&]() {
            std::vector<Bucket> bucket(((size_t)1 << c) - 1);
            for (;;) {
                size_t k = next.fetch_add(1);
                if (k >= n_windows) break;
                size_t i = n_windows - 1 - k;  // .rev()
                for (auto &b : bucket) b.kind = 0;
                for (size_t j = 0; j < n_pairs; j++) {
                    uint8_t bytes[32];
                    Fs::to_bytes(coeffs[j], bytes);  // to_raw_bytes per (window, scalar) (:26)
                    size_t seg = get_at(i, c, bytes);
                    if (seg != 0) bucket_add_assign(bucket[seg - 1], bases[j]);
                }
                Proj acc = Cv::proj_identity(), sum = Cv::proj_identity();
                for (size_t b = bucket.size(); b-- > 0;) {
                    sum = bucket_add(bucket[b], sum);
                    acc = Cv::add_proj(acc, sum);
                }
                for (size_t d = 0; d < c * i; d++) acc = Cv::double_proj(acc);
                window_acc[i] = acc;
            }
        };
        int nt = std::max(1, std::min<int>(threads, (int)n_windows));
        std::vector<std::thread> pool;
        for (int t = 1; t < nt; t++) pool.emplace_back(worker);
        worker();
        for (auto &t : pool) t.join();
        Proj total = Cv::proj_identity();
        for (size_t i = 0; i < n_windows; i++) total = Cv::add_proj(total, window_acc[i]);  // fold (:45-47)
        return total;
    }

    // nova/src/pedersen.rs:15-20: fold of sum + g_i * m_i, then .into() affine
    static Affine pedersen_commit(const Affine *g, size_t n_g, const Limbs *m, size_t n_m) {
        size_t n = std::min(n_g, n_m);
        Proj sum = Cv::proj_identity();
        for (size_t i = 0; i < n; i++) sum = Cv::add_proj(sum, Cv::scalar_point(Cv::to_extended(g[i]), m[i]));
        return Cv::to_affine(sum);
    }
};

inline Limbs ld4(const uint64_t *p) { return Limbs{p[0], p[1], p[2], p[3]}; }
inline void st4(uint64_t *p, const Limbs &l) { for (int i = 0; i < 4; i++) p[i] = l[i]; }

std::vector<Affine> load_affine(const uint64_t *xy, const uint8_t *inf, size_t n) {
    std::vector<Affine> v(n);
    for (size_t i = 0; i < n; i++) v[i] = Affine{ld4(xy + 8 * i), ld4(xy + 8 * i + 4), inf ? inf[i] != 0 : false};
    return v;
}
std::vector<Limbs> load_scalars(const uint64_t *s, size_t n) {
    std::vector<Limbs> v(n);
    for (size_t i = 0; i < n; i++) v[i] = ld4(s + 4 * i);
    return v;
}
void store_proj(uint64_t *out, const Proj &p) { st4(out, p.x); st4(out + 4, p.y); st4(out + 8, p.z); }
Proj load_proj(const uint64_t *in) { return Proj{ld4(in), ld4(in + 4), ld4(in + 8)}; }

template <class F> int field_op(int op, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    typedef Field<F> Fd;
    Limbs x = ld4(a), y = b ? ld4(b) : Limbs{0, 0, 0, 0}, r;
    switch (op) {
        case 0: r = Fd::add(x, y); break;
        case 1: r = Fd::sub(x, y); break;
        case 2: r = Fd::mul(x, y); break;
        case 3: r = Fd::square(x); break;
        case 4: r = Fd::dbl(x); break;
        case 5: r = Fd::neg(x); break;
        case 6: if (!Fd::invert(x, r)) return 1; break;
        case 7: r = Fd::montgomery_reduce(x); break;
        case 8: r = Fd::to_mont_form(x); break;
        case 9: { uint64_t w[8]; for (int i = 0; i < 4; i++) { w[i] = a[i]; w[4 + i] = b[i]; } r = Fd::from_u512(w); break; }
        default: return -1;
    }
    st4(out, r);
    return 0;
}

template <class C> void random_points(const uint8_t seed[16], size_t n, uint64_t *xy, uint64_t *scalars_out, int threads) {
    // macros/curve/weierstrass/group.rs:39-41: affine(G * Scalar::random(rng)); Scalar::random =
    // from_u512 of 8 next_u64 (represent.rs:80-103).  Scalars drawn serially (one stream), the
    // scalar multiplications are spread over threads.
    typedef Curve<C> Cv;
    typedef Field<typename C::Scalar> Fs;
    XorShift128 rng(seed);
    std::vector<Limbs> k(n);
    for (size_t i = 0; i < n; i++) {
        uint64_t w[8];
        for (int j = 0; j < 8; j++) w[j] = rng.next_u64();
        k[i] = Fs::from_u512(w);
        if (scalars_out) st4(scalars_out + 4 * i, k[i]);
    }
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (;;) {
            size_t i = next.fetch_add(64);
            if (i >= n) break;
            for (size_t j = i; j < std::min(n, i + 64); j++) {
                Affine a = Cv::to_affine(Cv::scalar_point(Cv::to_extended(Cv::generator()), k[j]));
                st4(xy + 8 * j, a.x);
                st4(xy + 8 * j + 4, a.y);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < std::max(1, threads); t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}

}  // namespace

#define DISPATCH_CURVE(curve, EXPR_G1, EXPR_GR) \
    do { if ((curve) == 0) { EXPR_G1; } else if ((curve) == 1) { EXPR_GR; } else return -1; } while (0)

extern "C" {

// field_id: 0 = bn254 Fq, 1 = bn254 Fr
int zko_field_op(int field_id, int op, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (field_id == 0) return field_op<FqParams>(op, a, b, out);
    if (field_id == 1) return field_op<FrParams>(op, a, b, out);
    return -1;
}

// curve: 0 = BN254 G1 (base Fq, scalar Fr), 1 = Grumpkin (base Fr, scalar Fq)
int zko_msm(int curve, const uint64_t *xy, const uint8_t *inf, size_t n_bases, const uint64_t *scalars, size_t n_scalars,
            int threads, uint64_t out[12]) {
    auto bases = load_affine(xy, inf, n_bases);
    auto sc = load_scalars(scalars, n_scalars);
    DISPATCH_CURVE(curve, store_proj(out, Msm<Bn254G1>::run(bases.data(), n_bases, sc.data(), n_scalars, threads)),
                   store_proj(out, Msm<Grumpkin>::run(bases.data(), n_bases, sc.data(), n_scalars, threads)));
    return 0;
}

// out = [x(4) y(4) inf(1 as u64)]
int zko_pedersen_commit(int curve, const uint64_t *xy, const uint8_t *inf, size_t n_g, const uint64_t *m, size_t n_m, uint64_t out[9]) {
    auto g = load_affine(xy, inf, n_g);
    auto sc = load_scalars(m, n_m);
    Affine r;
    DISPATCH_CURVE(curve, r = Msm<Bn254G1>::pedersen_commit(g.data(), n_g, sc.data(), n_m),
                   r = Msm<Grumpkin>::pedersen_commit(g.data(), n_g, sc.data(), n_m));
    st4(out, r.x); st4(out + 4, r.y); out[8] = r.inf;
    return 0;
}

int zko_to_affine(int curve, const uint64_t in[12], uint64_t out[9]) {
    Affine r;
    DISPATCH_CURVE(curve, r = Curve<Bn254G1>::to_affine(load_proj(in)), r = Curve<Grumpkin>::to_affine(load_proj(in)));
    st4(out, r.x); st4(out + 4, r.y); out[8] = r.inf;
    return 0;
}

// op: 0 add_proj(a,b), 1 double_proj(a), 2 add_mixed(affine b[x,y | inf in b[8]], proj a), 3 add_affine(a[0..8|inf a[8]], b[..]),
//     4 double_affine(a), 5 proj eq (out[0]), 6 is_on_curve(affine a) (out[0])
int zko_point_op(int curve, int op, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    auto aff = [](const uint64_t *p) { return Affine{ld4(p), ld4(p + 4), p[8] != 0}; };
#define BODY(C)                                                                         \
    {                                                                                   \
        typedef Curve<C> Cv;                                                            \
        switch (op) {                                                                   \
            case 0: store_proj(out, Cv::add_proj(load_proj(a), load_proj(b))); break;   \
            case 1: store_proj(out, Cv::double_proj(load_proj(a))); break;              \
            case 2: store_proj(out, Cv::add_mixed(aff(b), load_proj(a))); break;        \
            case 3: store_proj(out, Cv::add_affine(aff(a), aff(b))); break;             \
            case 4: store_proj(out, Cv::double_affine(aff(a))); break;                  \
            case 5: out[0] = Cv::eq(load_proj(a), load_proj(b)); break;                 \
            case 6: out[0] = Cv::is_on_curve(aff(a)); break;                            \
            default: return -1;                                                         \
        }                                                                               \
    }
    DISPATCH_CURVE(curve, BODY(Bn254G1), BODY(Grumpkin));
#undef BODY
    return 0;
}

// out = proj(12): point(a: proj 12) * scalar (Montgomery, scalar field of the curve)
int zko_scalar_point(int curve, const uint64_t a[12], const uint64_t scalar[4], uint64_t out[12]) {
    DISPATCH_CURVE(curve, store_proj(out, Curve<Bn254G1>::scalar_point(load_proj(a), ld4(scalar))),
                   store_proj(out, Curve<Grumpkin>::scalar_point(load_proj(a), ld4(scalar))));
    return 0;
}

int zko_generator(int curve, uint64_t out[8]) {
    Affine g;
    DISPATCH_CURVE(curve, g = Curve<Bn254G1>::generator(), g = Curve<Grumpkin>::generator());
    st4(out, g.x); st4(out + 4, g.y);
    return 0;
}

// Reference sampler (represent.rs:80-103) on the restated xorshift128 stream.
// field_id as in zko_field_op.  out: n x 4 u64 Montgomery.
int zko_random_field(int field_id, const uint8_t seed[16], size_t n, uint64_t *out) {
    XorShift128 rng(seed);
    for (size_t i = 0; i < n; i++) {
        uint64_t w[8];
        for (int j = 0; j < 8; j++) w[j] = rng.next_u64();
        if (field_id == 0) st4(out + 4 * i, Field<FqParams>::from_u512(w));
        else if (field_id == 1) st4(out + 4 * i, Field<FrParams>::from_u512(w));
        else return -1;
    }
    return 0;
}

// n random points G * k_i (k_i from the sampler above); optionally returns the k_i (Montgomery).
int zko_random_points(int curve, const uint8_t seed[16], size_t n, int threads, uint64_t *xy, uint64_t *scalars_out) {
    DISPATCH_CURVE(curve, random_points<Bn254G1>(seed, n, xy, scalars_out, threads),
                   random_points<Grumpkin>(seed, n, xy, scalars_out, threads));
    return 0;
}

int zko_xorshift_u64(const uint8_t seed[16], size_t n, uint64_t *out) {
    XorShift128 rng(seed);
    for (size_t i = 0; i < n; i++) out[i] = rng.next_u64();
    return 0;
}

// groth16/src/fft.rs on bn254 Fr.  op: 0 dft, 1 idft, 2 coset_dft, 3 coset_idft.  in: n_in x 4 u64 (Montgomery), out: 2^k x 4 u64,
// *n_out = number of meaningful elements (idft variants strip trailing zeros like Coefficients::new).
int zko_fft(size_t k, int op, const uint64_t *in, size_t n_in, uint64_t *out, size_t *n_out) {
    FrFft f(k);
    std::vector<Limbs> v = load_scalars(in, n_in), r;
    switch (op) {
        case 0: r = f.dft(v); break;
        case 1: r = f.idft(v); break;
        case 2: r = f.coset_dft(v); break;
        case 3: r = f.coset_idft(v); break;
        default: return -1;
    }
    *n_out = r.size();
    for (size_t i = 0; i < f.n; i++) st4(out + 4 * i, i < r.size() ? r[i] : Limbs{0, 0, 0, 0});
    return 0;
}
// prover.rs:36-47: coefficients of H from the R1CS evaluations (each m x 4 u64); out 2^k x 4 u64, *n_out after stripping
int zko_groth16_h(size_t k, const uint64_t *a, const uint64_t *b, const uint64_t *c, size_t m, uint64_t *out, size_t *n_out) {
    FrFft f(k);
    std::vector<Limbs> r = f.h_coefficients(load_scalars(a, m), load_scalars(b, m), load_scalars(c, m));
    *n_out = r.size();
    for (size_t i = 0; i < f.n; i++) st4(out + 4 * i, i < r.size() ? r[i] : Limbs{0, 0, 0, 0});
    return 0;
}

size_t zko_window_bits(size_t n_bases) { return window_bits(n_bases); }
size_t zko_get_at(size_t segment, size_t c, const uint8_t bytes[32]) { return get_at(segment, c, bytes); }

}  // extern "C"
