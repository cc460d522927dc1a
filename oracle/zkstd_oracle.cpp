// oracle/zkstd_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see zkstd_oracle.hpp header).
// extern "C" surface over the CPU restatement so tests / bench.py's cpu_baseline can drive it
// through ctypes.  Buffer formats are the same as include/kgr_msm.h so a test can hand the same
// arrays to both sides.
#include "zkstd_oracle.hpp"
#include <mutex>
#include <thread>
#include <atomic>
#include <algorithm>

using namespace zko;

namespace {

template <class C> struct Msm {
    typedef Curve<C> Cv;
    typedef Field<typename C::Scalar> Fs;
    typedef typename Cv::Affine Affine;
    typedef typename Cv::Proj Proj;

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
// coordinate <-> limbs: 4 per Fq/Fr element, 8 (c0 || c1) per Fq2 element
template <class E> struct Io;
template <> struct Io<Limbs> {
    static constexpr size_t N = 4;
    static Limbs ld(const uint64_t *p) { return ld4(p); }
    static void st(uint64_t *p, const Limbs &v) { st4(p, v); }
};
template <> struct Io<Fq2El> {
    static constexpr size_t N = 8;
    static Fq2El ld(const uint64_t *p) { return Fq2El{ld4(p), ld4(p + 4)}; }
    static void st(uint64_t *p, const Fq2El &v) { st4(p, v.c0); st4(p + 4, v.c1); }
};
template <class C> struct PtIo {
    typedef typename Curve<C>::El El;
    typedef typename Curve<C>::Affine Affine;
    typedef typename Curve<C>::Proj Proj;
    static constexpr size_t N = Io<El>::N;
    static std::vector<Affine> load_affine(const uint64_t *xy, const uint8_t *inf, size_t n) {
        std::vector<Affine> v(n);
        for (size_t i = 0; i < n; i++) v[i] = Affine{Io<El>::ld(xy + 2 * N * i), Io<El>::ld(xy + 2 * N * i + N), inf ? inf[i] != 0 : false};
        return v;
    }
    // x, y and a trailing is_infinity word
    static Affine affine1(const uint64_t *p) { return Affine{Io<El>::ld(p), Io<El>::ld(p + N), p[2 * N] != 0}; }
    static void store_affine1(uint64_t *out, const Affine &a) { Io<El>::st(out, a.x); Io<El>::st(out + N, a.y); out[2 * N] = a.inf; }
    static void store_proj(uint64_t *out, const Proj &p) { Io<El>::st(out, p.x); Io<El>::st(out + N, p.y); Io<El>::st(out + 2 * N, p.z); }
    static Proj load_proj(const uint64_t *in) { return Proj{Io<El>::ld(in), Io<El>::ld(in + N), Io<El>::ld(in + 2 * N)}; }
};
std::vector<Limbs> load_scalars(const uint64_t *s, size_t n) {
    std::vector<Limbs> v(n);
    for (size_t i = 0; i < n; i++) v[i] = ld4(s + 4 * i);
    return v;
}

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
                typename Cv::Affine a = Cv::to_affine(Cv::scalar_point(Cv::to_extended(Cv::generator()), k[j]));
                constexpr size_t N = PtIo<C>::N;
                Io<typename Cv::El>::st(xy + 2 * N * j, a.x);
                Io<typename Cv::El>::st(xy + 2 * N * j + N, a.y);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < std::max(1, threads); t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}


// ---- bench / test input support (NOT part of the reference restatement) ------------------------------------------------------------
// The synthetic inputs of bench.py are P_i = k_i * G with k_i = from_u512(8 words splitmix64(seed ^ splitmix64(8 i + j))): the stream the
// device generator (kgr_bases_generate_at) uses, restated here so that the CPU arm can build the SAME vectors without a GPU.
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
template <class F> void bench_scalars(uint64_t seed, uint64_t first, size_t n, uint64_t *out) {
    for (size_t i = 0; i < n; i++) {
        uint64_t w[8];
        for (int j = 0; j < 8; j++) w[j] = splitmix64(seed ^ splitmix64((first + i) * 8 + j));
        st4(out + 4 * i, Field<F>::from_u512(w));
    }
}
// k_i * G for many k_i (Montgomery scalars): byte-window table d * 2^(8 j) * G, at most 32 mixed additions per point with the reference's
// own formulas (add_mixed, weierstrass.rs:63-97) and to_affine.  Same group elements as scalar_point(G, k_i), hence the same affine limbs.
template <class C> void fixed_base_batch(const uint64_t *k, size_t n, int threads, uint64_t *xy) {
    typedef Curve<C> Cv;
    typedef Field<typename C::Scalar> Fs;
    constexpr size_t N = PtIo<C>::N;
    static std::vector<typename Cv::Affine> table;  // [32][255]
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (table.empty()) {
            table.resize(32 * 255);
            typename Cv::Proj base = Cv::to_extended(Cv::generator());
            for (int j = 0; j < 32; j++) {
                typename Cv::Proj acc = base;
                for (int d = 1; d <= 255; d++) {
                    table[j * 255 + d - 1] = Cv::to_affine(acc);
                    acc = Cv::add_proj(acc, base);
                }
                for (int b = 0; b < 8; b++) base = Cv::double_proj(base);
            }
        }
    }
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (;;) {
            size_t i0 = next.fetch_add(256);
            if (i0 >= n) break;
            const size_t i1 = std::min(n, i0 + 256), m = i1 - i0;
            typename Cv::Proj acc[256];
            typename Cv::El pre[256];
            typename Cv::El run = Cv::Fb::one();
            for (size_t t = 0; t < m; t++) {
                Limbs c = Fs::montgomery_reduce(ld4(k + 4 * (i0 + t)));
                acc[t] = Cv::proj_identity();
                for (int j = 0; j < 32; j++) {
                    unsigned d = (unsigned)(c[j / 8] >> (8 * (j % 8))) & 0xff;
                    if (d) acc[t] = Cv::add_mixed(table[j * 255 + d - 1], acc[t]);
                }
                pre[t] = run;  // product of the non-zero z before t
                if (!Cv::is_identity(acc[t])) run = Cv::Fb::mul(run, acc[t].z);
            }
            // one inversion per block (Montgomery's trick); x / z and y / z are the unique affine coordinates, as to_affine returns them
            typename Cv::El inv;
            Cv::Fb::invert(run, inv);
            for (size_t t = m; t-- > 0;) {
                typename Cv::El x = Cv::Fb::zero(), y = Cv::Fb::zero();  // identity (k_i = 0): the device generator's encoding (0, 0)
                if (!Cv::is_identity(acc[t])) {
                    typename Cv::El zi = Cv::Fb::mul(inv, pre[t]);
                    inv = Cv::Fb::mul(inv, acc[t].z);
                    x = Cv::Fb::mul(acc[t].x, zi);
                    y = Cv::Fb::mul(acc[t].y, zi);
                }
                Io<typename Cv::El>::st(xy + 2 * N * (i0 + t), x);
                Io<typename Cv::El>::st(xy + 2 * N * (i0 + t) + N, y);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < std::max(1, threads); t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}
// sum a_i * b_i in the field (Montgomery in and out): the discrete log of an MSM whose bases are k_i * G
template <class F> void field_dot(const uint64_t *a, const uint64_t *b, size_t n, int threads, uint64_t *out) {
    typedef Field<F> Fd;
    int T = std::max(1, threads);
    std::vector<Limbs> part(T, Limbs{0, 0, 0, 0});
    auto worker = [&](int t) {
        size_t lo = n * t / T, hi = n * (t + 1) / T;
        Limbs acc{0, 0, 0, 0};
        for (size_t i = lo; i < hi; i++) acc = Fd::add(acc, Fd::mul(ld4(a + 4 * i), ld4(b + 4 * i)));
        part[t] = acc;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < T; t++) pool.emplace_back(worker, t);
    worker(0);
    for (auto &t : pool) t.join();
    Limbs acc{0, 0, 0, 0};
    for (auto &p : part) acc = Fd::add(acc, p);
    st4(out, acc);
}

}  // namespace

// BODY is instantiated with C = the curve's parameter struct
#define DISPATCH_CURVE(curve, BODY)                                   \
    switch (curve) {                                                  \
        case 0: { typedef Bn254G1 C; BODY; } break;                   \
        case 1: { typedef Grumpkin C; BODY; } break;                  \
        case 2: { typedef Bn254G2 C; BODY; } break;                   \
        default: return -1;                                           \
    }

extern "C" {

// field_id: 0 = bn254 Fq, 1 = bn254 Fr
int zko_field_op(int field_id, int op, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (field_id == 0) return field_op<FqParams>(op, a, b, out);
    if (field_id == 1) return field_op<FrParams>(op, a, b, out);
    return -1;
}

// curve: 0 = BN254 G1 (base Fq, scalar Fr), 1 = Grumpkin (base Fr, scalar Fq), 2 = BN254 G2 (base Fq2, scalar Fr).
// Point buffers hold N = 4 limbs per coordinate (8 for G2): affine xy n x 2N, projective 3N, affine result 2N + 1.
int zko_msm(int curve, const uint64_t *xy, const uint8_t *inf, size_t n_bases, const uint64_t *scalars, size_t n_scalars,
            int threads, uint64_t *out) {
    auto sc = load_scalars(scalars, n_scalars);
    DISPATCH_CURVE(curve, {
        auto bases = PtIo<C>::load_affine(xy, inf, n_bases);
        PtIo<C>::store_proj(out, Msm<C>::run(bases.data(), n_bases, sc.data(), n_scalars, threads));
    });
    return 0;
}

// out = [x y inf(1 as u64)]
int zko_pedersen_commit(int curve, const uint64_t *xy, const uint8_t *inf, size_t n_g, const uint64_t *m, size_t n_m, uint64_t *out) {
    auto sc = load_scalars(m, n_m);
    DISPATCH_CURVE(curve, {
        auto g = PtIo<C>::load_affine(xy, inf, n_g);
        PtIo<C>::store_affine1(out, Msm<C>::pedersen_commit(g.data(), n_g, sc.data(), n_m));
    });
    return 0;
}

int zko_to_affine(int curve, const uint64_t *in, uint64_t *out) {
    DISPATCH_CURVE(curve, PtIo<C>::store_affine1(out, Curve<C>::to_affine(PtIo<C>::load_proj(in))));
    return 0;
}

// op: 0 add_proj(a,b), 1 double_proj(a), 2 add_mixed(affine b[x,y | inf], proj a), 3 add_affine(a[x,y | inf], b[..]),
//     4 double_affine(a), 5 proj eq (out[0]), 6 is_on_curve(affine a) (out[0])
int zko_point_op(int curve, int op, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    DISPATCH_CURVE(curve, {
        typedef Curve<C> Cv;
        typedef PtIo<C> P;
        switch (op) {
            case 0: P::store_proj(out, Cv::add_proj(P::load_proj(a), P::load_proj(b))); break;
            case 1: P::store_proj(out, Cv::double_proj(P::load_proj(a))); break;
            case 2: P::store_proj(out, Cv::add_mixed(P::affine1(b), P::load_proj(a))); break;
            case 3: P::store_proj(out, Cv::add_affine(P::affine1(a), P::affine1(b))); break;
            case 4: P::store_proj(out, Cv::double_affine(P::affine1(a))); break;
            case 5: out[0] = Cv::eq(P::load_proj(a), P::load_proj(b)); break;
            case 6: out[0] = Cv::is_on_curve(P::affine1(a)); break;
            default: return -1;
        }
    });
    return 0;
}

// out = proj: point(a: proj) * scalar (Montgomery, scalar field of the curve)
int zko_scalar_point(int curve, const uint64_t *a, const uint64_t scalar[4], uint64_t *out) {
    DISPATCH_CURVE(curve, PtIo<C>::store_proj(out, Curve<C>::scalar_point(PtIo<C>::load_proj(a), ld4(scalar))));
    return 0;
}

int zko_generator(int curve, uint64_t *out) {
    DISPATCH_CURVE(curve, {
        auto g = Curve<C>::generator();
        Io<typename Curve<C>::El>::st(out, g.x);
        Io<typename Curve<C>::El>::st(out + PtIo<C>::N, g.y);
    });
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
    DISPATCH_CURVE(curve, random_points<C>(seed, n, xy, scalars_out, threads));
    return 0;
}

// bench input support (see above): scalars of the curve's scalar field, fixed-base points, dot product in field 0 Fq / 1 Fr
int zko_bench_scalars(int curve, uint64_t seed, uint64_t first, size_t n, uint64_t *out) {
    DISPATCH_CURVE(curve, bench_scalars<typename C::Scalar>(seed, first, n, out));
    return 0;
}
int zko_fixed_base(int curve, const uint64_t *k, size_t n, int threads, uint64_t *xy) {
    DISPATCH_CURVE(curve, fixed_base_batch<C>(k, n, threads, xy));
    return 0;
}
int zko_field_dot(int field_id, const uint64_t *a, const uint64_t *b, size_t n, int threads, uint64_t *out) {
    if (field_id == 0) field_dot<FqParams>(a, b, n, threads, out);
    else if (field_id == 1) field_dot<FrParams>(a, b, n, threads, out);
    else return -1;
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

// zkstd/src/matrix.rs:36-48 SparseMatrix::prod over `field_id` (0 Fq, 1 Fr): out (m x 4) = M z
int zko_sparse_prod(int field_id, size_t m, const uint32_t *row_ptr, const uint32_t *cols, const uint64_t *coeffs, const uint64_t *z, size_t n_z, uint64_t *out) {
    CsrRef mat{row_ptr, cols, coeffs};
    std::vector<Limbs> zv = load_scalars(z, n_z), r;
    if (field_id == 0) r = NovaVec<FqParams>::prod(m, mat, zv);
    else if (field_id == 1) r = NovaVec<FrParams>::prod(m, mat, zv);
    else return -1;
    for (size_t i = 0; i < m; i++) st4(out + 4 * i, r[i]);
    return 0;
}
// nova/src/prover.rs:53-90 compute_cross_term; row_ptr / cols / coeffs: A, B, C
int zko_cross_term(int field_id, size_t m, const uint32_t *const row_ptr[3], const uint32_t *const cols[3], const uint64_t *const coeffs[3],
                   const uint64_t *z1, const uint64_t *z2, size_t n_z, uint64_t *out) {
    CsrRef a{row_ptr[0], cols[0], coeffs[0]}, b{row_ptr[1], cols[1], coeffs[1]}, c{row_ptr[2], cols[2], coeffs[2]};
    std::vector<Limbs> v1 = load_scalars(z1, n_z), v2 = load_scalars(z2, n_z), r;
    if (field_id == 0) r = NovaVec<FqParams>::cross_term(m, a, b, c, v1, v2);
    else if (field_id == 1) r = NovaVec<FrParams>::cross_term(m, a, b, c, v1, v2);
    else return -1;
    for (size_t i = 0; i < m; i++) st4(out + 4 * i, r[i]);
    return 0;
}
// nova/src/relaxed_r1cs/witness.rs:67-68: out = a + b * r
int zko_vec_fold(int field_id, const uint64_t *a, const uint64_t *b, const uint64_t r[4], size_t n, uint64_t *out) {
    std::vector<Limbs> av = load_scalars(a, n), bv = load_scalars(b, n), o;
    if (field_id == 0) o = NovaVec<FqParams>::fold(av, bv, ld4(r));
    else if (field_id == 1) o = NovaVec<FrParams>::fold(av, bv, ld4(r));
    else return -1;
    for (size_t i = 0; i < n; i++) st4(out + 4 * i, o[i]);
    return 0;
}

size_t zko_window_bits(size_t n_bases) { return window_bits(n_bases); }
size_t zko_get_at(size_t segment, size_t c, const uint8_t bytes[32]) { return get_at(segment, c, bytes); }

}  // extern "C"
