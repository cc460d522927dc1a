// affine_kernels.cuh — batched-affine tree levels in front of the XYZZ bucket accumulation (round 2; kgr_set_param("affine_levels", r)).
//
// The reference's inner operation is a mixed addition per (scalar, window) (groth16/src/msm.rs:25-33 -> weierstrass.rs:63-97, 11 field
// multiplications; 10 in the XYZZ form k_accumulate uses).  An AFFINE addition costs 1M + 1S + 1M once the inverse of its denominator is
// known, and Montgomery's trick shares one inversion between any number of independent denominators at 3 multiplications each, so a
// level of pairwise sums inside every bucket costs ~6.5 multiplications per addition instead of 10.
//
// Level l -> l + 1, for every bucket g with len_l nodes at positions [off_l[g], off_l[g + 1]):
//     node_{l+1}[off_{l+1}[g] + j] = node_l[off_l[g] + 2 j] + node_l[off_l[g] + 2 j + 1]      (the odd last node is copied)
// with off_{l+1} = exclusive scan of ceil(len_l / 2).  Level 0 is the sorted entry list itself (base index | sign << 31).  After r levels
// a bucket holds ceil(len / 2^r) affine nodes, which k_accumulate sums exactly as it sums base points (entries == nullptr: node = position).
//
// Round 1's version of the idea (body_accumulate_affine, msm_kernels.cuh) kept each thread's tree in a private slice of global scratch and
// paid one inversion per THREAD: 10x the DRAM traffic and an 82-multiplication inversion per ~100 additions (profiles/r01_affine.md).
// A first round-2 version did phase 1, a CTA-wide inversion and phase 3 in ONE kernel: correct, but the single-thread inversion and the
// product tree ran with 1 .. 64 of 128 threads active (21 of 32 lanes active on average, ncu) and it lost to the XYZZ kernel
// (profiles/r02_affine.md).  Now a level is three kernels, each with every lane busy:
//   k_affine_den   thread = AFF_K consecutive OUTPUT nodes: chord / tangent denominators front to back, prefix products to a coalesced
//                  scratch array, the thread's total to tot[]
//   k_batch_inv    tot[] inverted in place: each CTA shares ONE safegcd inversion between 2048 values (per-thread prefixes, product
//                  tree in shared memory, inverses pushed back down) — 1 inversion per 32 K additions
//   k_affine_add   back to front: 1 / den_j = inv * prefix_j, lambda, x3, y3 (5 multiplications), output node stored
// All inputs and outputs are whole 64-byte points at consecutive addresses per thread.
#pragma once
#include "msm_kernels.cuh"

namespace kgr {

constexpr int AFF_TPB = 128;  // threads per CTA of the level kernels
constexpr int AFF_K = 16;     // output nodes per thread
constexpr int INV_K = 16;     // values per thread of k_batch_inv (one inversion per AFF_TPB * INV_K values)

// level-l nodes: the sorted entries over the base array (level 0) or an array of affine points (level >= 1)
template <class C> struct AffLevelIn {
    const AffinePt<C> *bases;
    const uint32_t *entries;  // nullptr: nodes are addressed directly
    // level 0, optional: the x coordinates of the bases alone, one 32-byte element per point (k_extract_x).  k_affine_den needs nothing else, and
    // a gather from the dense array moves half the bytes of a gather of the x half of a 64-byte point (2^21 - 2^23 points: accumulate phase - 2 %)
    const typename C::Elem *xs = nullptr;
    __device__ __forceinline__ const AffinePt<C> *addr(uint32_t pos, bool &neg) const {
        if (entries) {
            uint32_t ent = entries[pos];
            neg = (ent >> 31) != 0;
            return bases + (ent & 0x7fffffffu);
        }
        neg = false;
        return bases + pos;
    }
    __device__ __forceinline__ AffinePt<C> load(uint32_t pos) const {
        bool neg;
        const AffinePt<C> *p = addr(pos, neg);
        AffinePt<C> r = load_affine(p, 0);
        r.y = fp_cneg(r.y, neg);
        return r;
    }
};

// shared-memory field elements, word-major (word k of slot s at [k * stride + s]): conflict-free for consecutive slots
template <class E> __device__ __forceinline__ void sm_put_el(uint32_t *sm, int stride, int slot, const E &v) {
#pragma unroll
    for (int k = 0; k < El<E>::WORDS; k++) sm[k * stride + slot] = El<E>::word(v, k);
}
template <class E> __device__ __forceinline__ E sm_get_el(const uint32_t *sm, int stride, int slot) {
    E v;
#pragma unroll
    for (int k = 0; k < El<E>::WORDS; k++) El<E>::word(v, k) = sm[k * stride + slot];
    return v;
}

// xs[i] = pts[i].x
template <class C> __global__ void __launch_bounds__(256) k_extract_x(const AffinePt<C> *pts, uint32_t n, typename C::Elem *xs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename C::Elem x;
    el_load(x, &pts[i].x);
    el_store(&xs[i], x);
}

// cnt[g] = ceil(len_l(g) / 2) for g < G, cnt[G] = 0 (so that the exclusive scan over G + 1 elements ends with the total)
static __global__ void __launch_bounds__(256) k_affine_counts(const uint32_t *off_in, uint32_t G, uint32_t *cnt) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) cnt[g] = (off_in[g + 1] - off_in[g] + 1) >> 1;
    else if (g == G) cnt[g] = 0;
}

// Where output node q of a level comes from: bucket g with off_out[g] <= q < off_out[g + 1] and its input range.
struct AffWalk {
    uint32_t g, g_lo, g_hi, in_lo, in_hi;
    __device__ __forceinline__ void start(const uint32_t *off_in, const uint32_t *off_out, uint32_t G, uint32_t q) {
        g = bucket_of_position(off_out, G, q);
        g_lo = off_out[g];
        g_hi = off_out[g + 1];
        in_lo = off_in[g];
        in_hi = off_in[g + 1];
    }
    __device__ __forceinline__ void forward(const uint32_t *off_in, const uint32_t *off_out, uint32_t q) {
        while (q >= g_hi) {
            g++;
            g_lo = g_hi;
            g_hi = off_out[g + 1];
            in_lo = in_hi;
            in_hi = off_in[g + 1];
        }
    }
    __device__ __forceinline__ void backward(const uint32_t *off_in, const uint32_t *off_out, uint32_t q) {
        while (q < g_lo) {  // previous non-empty bucket (an empty one has off_out[g] == off_out[g + 1] > q)
            g--;
            g_hi = g_lo;
            g_lo = off_out[g];
            in_hi = in_lo;
            in_lo = off_in[g];
        }
    }
};

// scratch layout: word w of (slot j, thread t) at ((j * EW + w) * NT + t): every warp access is 128 consecutive bytes
template <class C>
__global__ void __launch_bounds__(AFF_TPB, 4) k_affine_den(AffLevelIn<C> in, const uint32_t *off_in, const uint32_t *off_out, uint32_t G, uint32_t NT, uint32_t *pre,
                                                            uint32_t *tot) {
    typedef typename C::Elem E;
    constexpr int EW = El<E>::WORDS;
    const uint32_t t = blockIdx.x * AFF_TPB + threadIdx.x;
    const uint32_t total = off_out[G];
    const uint64_t q0_64 = (uint64_t)t * AFF_K;
    E run = El<E>::one();
    if (q0_64 < total) {
        const uint32_t q0 = (uint32_t)q0_64, nq = min((uint32_t)AFF_K, total - q0);
        AffWalk wk;
        wk.start(off_in, off_out, G, q0);
        for (uint32_t j = 0; j < nq; j++) {
            const uint32_t q = q0 + j;
            wk.forward(off_in, off_out, q);
            const uint32_t p0 = wk.in_lo + 2 * (q - wk.g_lo);
            E den = El<E>::one();
            if (p0 + 1 < wk.in_hi) {
                bool na, nb;
                const AffinePt<C> *pa = in.addr(p0, na), *pb = in.addr(p0 + 1, nb);
                E xa, xb;   // the chord denominator needs the x coordinates only
                if (in.xs) {
                    el_load(xa, in.xs + (pa - in.bases));
                    el_load(xb, in.xs + (pb - in.bases));
                } else {
                    xa = load_x(pa);
                    xb = load_x(pb);
                }
                den = fp_sub(xb, xa);
                if (fp_is_zero(den) || fp_is_zero(xa) || fp_is_zero(xb)) {  // equal x, or x = 0 (maybe the identity encoding (0, 0))
                    AffinePt<C> a = in.load(p0), b = in.load(p0 + 1);
                    (void)pair_case(a, b, den);
                }
            }
#pragma unroll
            for (int w = 0; w < EW; w++) pre[((size_t)j * EW + w) * NT + t] = El<E>::word(run, w);
            run = fp_mul(run, den);
        }
    }
    if (t < NT) {
#pragma unroll
        for (int w = 0; w < EW; w++) tot[(size_t)w * NT + t] = El<E>::word(run, w);
    }
}

// v[0 .. n) (word-major, stride NT) inverted in place; all values are non-zero.  One inversion per CTA (AFF_TPB * INV_K values).
// The per-thread prefix products go to `pfx` (same layout as v) rather than to shared memory: the CTA then holds 8 KB of shared memory instead
// of 40, so that enough CTAs are resident per SM to fill the ~100 dependent multiplications of the product tree and the single-thread
// inversion with the other CTAs' prefix / back-substitution work (r02: 15.8 of 32 lanes active and 58 % multiplier utilisation before).
template <class E> __global__ void __launch_bounds__(AFF_TPB) k_batch_inv(uint32_t *v, uint32_t n, uint32_t NT, uint32_t *pfx, const uint32_t *total_out) {
    constexpr int EW = El<E>::WORDS;
    __shared__ uint32_t s_tree[EW * 2 * AFF_TPB];
    const int tid = threadIdx.x;
    const uint32_t base = blockIdx.x * (AFF_TPB * INV_K);
    // the grid is sized for the host's bound on the output nodes; only the threads of k_affine_den / k_affine_add that own nodes use their total
    // (skewed scalars: a fraction of the bound — the other CTAs leave at once)
    n = min(n, (*total_out + (uint32_t)AFF_K - 1) / (uint32_t)AFF_K);
    if (base >= n) return;
    E run = El<E>::one();
#pragma unroll 1
    for (int i = 0; i < INV_K; i++) {
        const uint32_t idx = base + i * AFF_TPB + tid;
        if (idx < n) {
            E x;
#pragma unroll
            for (int w = 0; w < EW; w++) {
                El<E>::word(x, w) = v[(size_t)w * NT + idx];
                pfx[(size_t)w * NT + idx] = El<E>::word(run, w);
            }
            run = fp_mul(run, x);
        }
    }
    sm_put_el(s_tree, 2 * AFF_TPB, AFF_TPB + tid, run);
    __syncthreads();
    for (int s = AFF_TPB / 2; s >= 1; s >>= 1) {
        if (tid < s) {
            E l = sm_get_el<E>(s_tree, 2 * AFF_TPB, 2 * (s + tid)), r = sm_get_el<E>(s_tree, 2 * AFF_TPB, 2 * (s + tid) + 1);
            sm_put_el(s_tree, 2 * AFF_TPB, s + tid, fp_mul(l, r));
        }
        __syncthreads();
    }
    if (tid == 0) sm_put_el(s_tree, 2 * AFF_TPB, 1, fp_inv_fast(sm_get_el<E>(s_tree, 2 * AFF_TPB, 1)));
    __syncthreads();
    for (int s = 1; s < AFF_TPB; s <<= 1) {
        if (tid < s) {
            const int p = s + tid;
            E ip = sm_get_el<E>(s_tree, 2 * AFF_TPB, p);
            E l = sm_get_el<E>(s_tree, 2 * AFF_TPB, 2 * p), r = sm_get_el<E>(s_tree, 2 * AFF_TPB, 2 * p + 1);
            sm_put_el(s_tree, 2 * AFF_TPB, 2 * p, fp_mul(ip, r));
            sm_put_el(s_tree, 2 * AFF_TPB, 2 * p + 1, fp_mul(ip, l));
        }
        __syncthreads();
    }
    E inv_run = sm_get_el<E>(s_tree, 2 * AFF_TPB, AFF_TPB + tid);
#pragma unroll 1
    for (int i = INV_K; i-- > 0;) {
        const uint32_t idx = base + i * AFF_TPB + tid;
        if (idx < n) {
            E x, pj;
#pragma unroll
            for (int w = 0; w < EW; w++) {
                El<E>::word(x, w) = v[(size_t)w * NT + idx];
                El<E>::word(pj, w) = pfx[(size_t)w * NT + idx];
            }
            E xi = fp_mul(inv_run, pj);
            inv_run = fp_mul(inv_run, x);
#pragma unroll
            for (int w = 0; w < EW; w++) v[(size_t)w * NT + idx] = El<E>::word(xi, w);
        }
    }
}

template <class C>
__global__ void __launch_bounds__(AFF_TPB, 4) k_affine_add(AffLevelIn<C> in, const uint32_t *off_in, const uint32_t *off_out, uint32_t G, uint32_t NT, const uint32_t *pre,
                                                            const uint32_t *tot, AffinePt<C> *out) {
    typedef typename C::Elem E;
    constexpr int EW = El<E>::WORDS;
    const uint32_t t = blockIdx.x * AFF_TPB + threadIdx.x;
    const uint32_t total = off_out[G];
    const uint64_t q0_64 = (uint64_t)t * AFF_K;
    if (q0_64 >= total) return;
    const uint32_t q0 = (uint32_t)q0_64, nq = min((uint32_t)AFF_K, total - q0);
    E inv_run;  // 1 / (product of this thread's denominators)
#pragma unroll
    for (int w = 0; w < EW; w++) El<E>::word(inv_run, w) = tot[(size_t)w * NT + t];
    AffWalk wk;
    wk.start(off_in, off_out, G, q0 + nq - 1);
    for (uint32_t j = nq; j-- > 0;) {
        const uint32_t q = q0 + j;
        wk.backward(off_in, off_out, q);
        const uint32_t p0 = wk.in_lo + 2 * (q - wk.g_lo);
        AffinePt<C> a = in.load(p0), r = a;
        if (p0 + 1 < wk.in_hi) {
            AffinePt<C> b = in.load(p0 + 1);
            E den, pj;
            const int code = pair_case(a, b, den);
#pragma unroll
            for (int w = 0; w < EW; w++) El<E>::word(pj, w) = pre[((size_t)j * EW + w) * NT + t];
            E den_inv = fp_mul(inv_run, pj);
            inv_run = fp_mul(inv_run, den);
            r = pair_sum(code, a, b, den_inv);
        }
        el_store(&out[q].x, r.x);
        el_store(&out[q].y, r.y);
    }
}

}  // namespace kgr
