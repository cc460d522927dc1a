// kernels_curve.cuh — __global__ wrappers of the per-thread bodies (msm_kernels.cuh), templated on the curve.
// Included only by launch_impl.cuh, which is compiled once per curve (kernels_g1.cu, kernels_grumpkin.cu, kernels_g2_*.cu).
#pragma once
#include <cuda_runtime.h>

#include "launch.cuh"
#include "msm_kernels.cuh"

namespace kgr {

// resident CTAs per SM the accumulate kernels are compiled for: 4 (128 registers) with 8-word coordinates, 2 (255 registers) for G2,
// whose XYZZ accumulator alone is 64 registers
template <class C> constexpr int xyzz_words() { return (int)(sizeof(XyzzPt<C>) / 4); }
template <class C> constexpr int acc_min_blocks() { return El<typename C::Elem>::WORDS == 8 ? 4 : 2; }


template <class C>
__global__ void __launch_bounds__(TPB_SCALAR) k_count(MsmShape sh, const uint32_t *scalars, int is_mont, uint32_t *counts, uint32_t *digits) {
    body_count<C>(blockIdx.x * blockDim.x + threadIdx.x, sh, scalars, is_mont, counts, digits);
}
template <class C>
__global__ void __launch_bounds__(TPB_SCALAR) k_fill_window(MsmShape sh, const uint32_t *digits, uint32_t *counts, const uint32_t *offsets, uint32_t *entries) {
    body_fill_window<C>(blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y, sh, digits, counts, offsets, entries);
}
template <class C>
__global__ void __launch_bounds__(TPB_SCALAR) k_fill(MsmShape sh, const uint32_t *scalars, int is_mont, uint32_t *counts, const uint32_t *offsets,
                                                     uint32_t *entries) {
    body_fill<C>(blockIdx.x * blockDim.x + threadIdx.x, sh, scalars, is_mont, counts, offsets, entries);
}
template <class C>
__global__ void __launch_bounds__(TPB_ACC, acc_min_blocks<C>()) k_accumulate(MsmShape sh, const AffinePt<C> *bases, const uint32_t *offsets, const uint32_t *entries,
                                                        XyzzPt<C> *bucket_acc, XyzzPt<C> *head, XyzzPt<C> *tail, uint32_t *tail_bucket) {
    body_accumulate<C>(blockIdx.x * blockDim.x + threadIdx.x, sh, bases, offsets, entries, bucket_acc, head, tail, tail_bucket);
}
// Hot buckets (cut into more than FIXUP_INLINE_MAX pieces: skewed scalars, a thin top window) are summed by k_fixup_long.  A bucket of P pieces
// becomes S = ceil(P / FL_SLICE) <= FL_MAX_SLICES worklist records (bucket, slice | S << 8, index of its first record, -), one CTA each, so that a
// bucket holding a large part of the input (Nova witnesses: half the scalars equal) is summed by many CTAs instead of by the 128 lanes of one
// (2^20 points, all scalars 0 / 1: 8192 pieces, 64 dependent additions per lane before).  wl: 4 words per record.
constexpr uint32_t FL_SLICE = 512, FL_MAX_SLICES = 64;
KGR_D void push_hot_bucket(uint32_t g, uint32_t pieces, uint32_t *wl, uint32_t *wl_len) {
    uint32_t S = (pieces + FL_SLICE - 1) / FL_SLICE;
    if (S > FL_MAX_SLICES) S = FL_MAX_SLICES;
    const uint32_t base = atomic_add_u32(wl_len, S);
    for (uint32_t k = 0; k < S; k++) {
        wl[4 * (size_t)(base + k)] = g;
        wl[4 * (size_t)(base + k) + 1] = k | (S << 8);
        wl[4 * (size_t)(base + k) + 2] = base;
    }
}
// body_fixup with FOUR lanes per chunk (xyzz_add_quad): the up to FIXUP_INLINE_MAX dependent additions of a cut bucket are the longest chain
// between the accumulation and the reduction (6 x ~7 us for a lone thread; 20 us each over Fq2).  Same reads, same sums, same result.
template <class C>
__global__ void __launch_bounds__(TPB_RED) k_fixup(MsmShape sh, const uint32_t *offsets, XyzzPt<C> *bucket_acc, const XyzzPt<C> *head,
                                                   const XyzzPt<C> *tail, const uint32_t *tail_bucket, uint32_t *worklist, uint32_t *worklist_len) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, t = tid >> 2;
    const uint32_t M = offsets[sh.G], L = eff_chunk_len(sh, M);
    if ((uint64_t)t * L >= M) return;  // the same for the four lanes of a quad
    const uint32_t g = tail_bucket[t];
    if (g == NO_DIGIT) return;
    const uint32_t t1 = (offsets[g + 1] - 1) / L;
    if (t1 - t > FIXUP_INLINE_MAX) {
        if ((tid & 3) == 0) push_hot_bucket(g, t1 - t + 1, worklist, worklist_len);
        return;
    }
    XyzzPt<C> acc = tail[t];
    for (uint32_t u = t + 1; u <= t1; u++) {
        XyzzPt<C> h = head[u];
        xyzz_add_quad(acc, h);
    }
    if ((tid & 3) == 0) store_xyzz(&bucket_acc[g], acc);
}

// The same sums with four lanes per BUCKET: when the chunks outnumber the buckets (small inputs: every bucket is cut into several chunks), a
// quad per chunk leaves most lanes of a warp idle while the warp still pays for every multiplication (2^15 points: 82 us); a quad per bucket
// keeps the lanes dense.  Bucket g starts in chunk t0 (its first piece is tail[t0]) and continues through head[t0 + 1 .. t1].
template <class C>
__global__ void __launch_bounds__(TPB_RED) k_fixup_buckets(MsmShape sh, const uint32_t *offsets, XyzzPt<C> *bucket_acc, const XyzzPt<C> *head,
                                                           const XyzzPt<C> *tail, uint32_t *worklist, uint32_t *worklist_len) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, g = tid >> 2;
    if (g >= sh.G) return;
    const uint32_t lo = offsets[g], hi = offsets[g + 1];
    if (lo == hi) return;
    const uint32_t L = eff_chunk_len(sh, offsets[sh.G]);
    const uint32_t t0 = lo / L, t1 = (hi - 1) / L;
    if (t0 == t1) return;  // inside one chunk: the accumulate kernel wrote the bucket itself
    if (t1 - t0 > FIXUP_INLINE_MAX) {
        if ((tid & 3) == 0) push_hot_bucket(g, t1 - t0 + 1, worklist, worklist_len);
        return;
    }
    XyzzPt<C> acc = tail[t0];
    for (uint32_t u = t0 + 1; u <= t1; u++) {
        XyzzPt<C> h = head[u];
        xyzz_add_quad(acc, h);
    }
    if ((tid & 3) == 0) store_xyzz(&bucket_acc[g], acc);
}

// XYZZ points in shared memory, word-major (word k of thread t at sm[k * TPB + t]): conflict-free.
template <class C> __device__ __forceinline__ void sm_put(uint32_t *sm, int t, const XyzzPt<C> &p) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(&p);
#pragma unroll
    for (int k = 0; k < xyzz_words<C>(); k++) sm[k * TPB_TREE + t] = w[k];
}
template <class C> __device__ __forceinline__ XyzzPt<C> sm_get(const uint32_t *sm, int t) {
    XyzzPt<C> p;
    uint32_t *w = reinterpret_cast<uint32_t *>(&p);
#pragma unroll
    for (int k = 0; k < xyzz_words<C>(); k++) w[k] = sm[k * TPB_TREE + t];
    return p;
}
// Sum of the first 2 * s0 points of the shared-memory array (s0 a power of two <= TPB_TREE / 4), left in slot 0: level s adds slot qd + s to
// slot qd with the four lanes of quad qd working on one addition (xyzz_add_quad), log2(2 s0) quad-addition latencies.  Every thread of the
// CTA calls it (barriers inside); sm must be complete and visible (barrier before the call).
template <class C> __device__ __forceinline__ void quad_tree_sum(uint32_t *sm, int s0) {
    const int t = threadIdx.x, qd = t >> 2;
    for (int s = s0; s > 0; s >>= 1) {
        if (qd < s) {
            XyzzPt<C> a = sm_get<C>(sm, qd), b = sm_get<C>(sm, qd + s);
            xyzz_add_quad(a, b);
            __syncwarp(0xFu << (t & 28));  // the quad's reads of slot qd are complete (its shuffles order them already; this names it)
            if ((t & 3) == 0) sm_put<C>(sm, qd, a);
        }
        __syncthreads();
    }
}
// Sum of the TPB_TREE per-thread points of a CTA, returned in thread 0: one level with a thread per addition, then the quad tree.
template <class C> __device__ __forceinline__ XyzzPt<C> block_tree_sum(XyzzPt<C> v, uint32_t *sm) {
    const int t = threadIdx.x;
    sm_put<C>(sm, t, v);
    __syncthreads();
    for (int s = TPB_TREE / 2; s > TPB_TREE / 4; s >>= 1) {
        if (t < s) {
            XyzzPt<C> o = sm_get<C>(sm, t + s);
            xyzz_add(v, o);
            sm_put<C>(sm, t, v);
        }
        __syncthreads();
    }
    quad_tree_sum<C>(sm, TPB_TREE / 4);
    return sm_get<C>(sm, 0);
}
// One CTA per worklist record (grid-stride): the CTA sums its slice of the bucket's pieces (piece 0 is tail[t0], piece k >= 1 is head[t0 + k]);
// the CTA that finishes a multi-slice bucket LAST (counter + __threadfence) adds the S slice sums.  counter[] is zero on entry and is left zero.
template <class C> __device__ __forceinline__ XyzzPt<C> load_xyzz_cg(const XyzzPt<C> *p) {  // written by other CTAs of this launch: not through the read-only path
    XyzzPt<C> r;
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint32_t *w = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int k = 0; k < xyzz_words<C>() / 4; k++) {
        uint4 a = __ldcg(q + k);
        w[4 * k] = a.x; w[4 * k + 1] = a.y; w[4 * k + 2] = a.z; w[4 * k + 3] = a.w;
    }
    return r;
}
template <class C>
__global__ void __launch_bounds__(TPB_TREE) k_fixup_long(MsmShape sh, const uint32_t *offsets, XyzzPt<C> *bucket_acc, const XyzzPt<C> *head,
                                                         const XyzzPt<C> *tail, const uint32_t *worklist, const uint32_t *worklist_len, XyzzPt<C> *partial,
                                                         uint32_t *counter) {
    __shared__ uint32_t sm[xyzz_words<C>() * TPB_TREE];
    __shared__ uint32_t last_flag;
    static_assert(FL_MAX_SLICES <= TPB_TREE, "one lane per slice sum");
    const uint32_t n = *worklist_len;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint32_t g = worklist[4 * (size_t)i], ss = worklist[4 * (size_t)i + 1], base = worklist[4 * (size_t)i + 2];
        const uint32_t sl = ss & 0xffu, S = ss >> 8;
        const uint32_t lo = offsets[g], hi = offsets[g + 1];
        const uint32_t L = eff_chunk_len(sh, offsets[sh.G]);
        const uint32_t t0 = lo / L, t1 = (hi - 1) / L, P = t1 - t0 + 1;
        const uint32_t per = (P + S - 1) / S, first = sl * per, last = min(P, first + per);
        XyzzPt<C> v = xyzz_identity<C>();
        for (uint32_t p = first + threadIdx.x; p < last; p += TPB_TREE) {
            XyzzPt<C> x = p == 0 ? tail[t0] : head[t0 + p];
            xyzz_add(v, x);
        }
        v = block_tree_sum<C>(v, sm);
        if (S == 1) {
            if (threadIdx.x == 0) store_xyzz(&bucket_acc[g], v);
        } else {
            if (threadIdx.x == 0) {
                store_xyzz(&partial[i], v);
                __threadfence();
                last_flag = atomicAdd(&counter[base], 1u) == S - 1;
            }
            __syncthreads();
            if (last_flag) {
                __threadfence();
                XyzzPt<C> q = threadIdx.x < S ? load_xyzz_cg<C>(&partial[base + threadIdx.x]) : xyzz_identity<C>();
                __syncthreads();
                q = block_tree_sum<C>(q, sm);
                if (threadIdx.x == 0) {
                    store_xyzz(&bucket_acc[g], q);
                    counter[base] = 0;
                }
            }
        }
        __syncthreads();
    }
}
template <class C>
__global__ void __launch_bounds__(TPB_RED) k_bucket_merge(uint32_t G, const uint32_t *piece_offsets, const XyzzPt<C> *piece_acc, XyzzPt<C> *bucket_acc) {
    body_bucket_merge<C>(blockIdx.x * blockDim.x + threadIdx.x, G, piece_offsets, piece_acc, bucket_acc);
}
template <class C>
__global__ void __launch_bounds__(TPB_RED) k_reduce(uint32_t n_windows, uint32_t cnt_in, uint32_t K, uint32_t m_log2, const XyzzPt<C> *in_s,
                                                    const XyzzPt<C> *in_a, XyzzPt<C> *out_s, XyzzPt<C> *out_a, const uint32_t *bucket_offsets) {
    body_reduce<C>(blockIdx.x * blockDim.x + threadIdx.x, n_windows, cnt_in, K, m_log2, in_s, in_a, out_s, out_a, bucket_offsets);
}
template <class C>
__global__ void __launch_bounds__(TPB_RED) k_weight(uint32_t n_windows, uint32_t cnt, uint32_t m_log2, const XyzzPt<C> *in_s, const XyzzPt<C> *in_a,
                                                    XyzzPt<C> *out) {
    body_weight<C>(blockIdx.x * blockDim.x + threadIdx.x, n_windows, cnt, m_log2, in_s, in_a, out);
}
// grid (ceil(cnt_in / TPB_TREE), windows): out[w][block] = sum of in[w][block * TPB_TREE ...]
template <class C> __global__ void __launch_bounds__(TPB_TREE) k_tree_sum(const XyzzPt<C> *in, uint32_t cnt_in, XyzzPt<C> *out) {
    __shared__ uint32_t sm[xyzz_words<C>() * TPB_TREE];
    uint32_t i = blockIdx.x * TPB_TREE + threadIdx.x, w = blockIdx.y;
    XyzzPt<C> v = (i < cnt_in) ? in[(size_t)w * cnt_in + i] : xyzz_identity<C>();
    v = block_tree_sum<C>(v, sm);
    if (threadIdx.x == 0) store_xyzz(&out[(size_t)w * gridDim.x + blockIdx.x], v);
}
// ---- fold reduce ----------------------------------------------------------------------------------------
// Window sum S = sum_i (i + 1) * X_i over B = 2^nb buckets, written as S = T0 + sum_b 2^b * V_b with T0 = sum of all buckets
// and V_b = sum of the buckets whose index has bit b set.  Folding the array in half (Y_i = X_i + X_{i+m}) keeps T0 and
// every lower V_b and exposes V_{log2 m} as the plain sum of the upper half, so nb fully parallel fold levels (B adds in
// total) plus plain sums of the upper halves (B adds) replace the serial running sums; every level is one add deep.
// Level l (m = B >> l) reads Y^(l-1) (the buckets for l = 1) and writes Y^(l) at offset B - 2m of the window's row in F.
template <class C>
__global__ void __launch_bounds__(TPB_RED) k_fold(const XyzzPt<C> *buckets, XyzzPt<C> *F, uint32_t B, uint32_t l, uint32_t n_windows, const uint32_t *bucket_offsets) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, m = B >> l;
    if (t >= n_windows * m) return;
    body_fold<C>(t / m, t % m, l, B, buckets, F, bucket_offsets);
}
// The same level with four lanes per addition (xyzz_add_quad), for the levels that no longer fill the GPU with a thread per addition: there a
// launch lasts one addition of a lone warp (15 - 17 us per level at 2^20 points), a quad needs a third of it.  l >= 2 (the input is a fold level).
template <class C>
__global__ void __launch_bounds__(TPB_RED) k_fold_quad(XyzzPt<C> *F, uint32_t B, uint32_t l, uint32_t n_windows) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, t = tid >> 2, m = B >> l;
    if (t >= n_windows * m) return;
    const uint32_t w = t / m, i = t % m;
    const XyzzPt<C> *in = F + (size_t)w * B + fold_level_offset(B, l - 1);
    XyzzPt<C> a = in[i], b = in[i + m];
    xyzz_add_quad(a, b);
    if ((tid & 3) == 0) store_xyzz(&F[(size_t)w * B + fold_level_offset(B, l) + i], a);
}
// Fold levels l_first .. nb of one window in one CTA (m = B >> l <= TPB_TAIL there): the deep levels are one add each and purely
// latency-bound, so they are not worth a launch apiece.  Same reads and writes as k_fold; a level's output is the next level's input,
// ordered by the CTA barrier.
constexpr int TPB_TAIL = 256;
constexpr uint32_t VSUM_ELEMS = 8;  // 8 serial additions + a 7-level tree per CTA: one wave of CTAs at B = 2^14 (2: four times the CTAs and their trees, reduce 0.46 -> 0.55 ms at 2^20)
constexpr uint32_t VSUM_ELEMS_LATE = 1;  // the levels summed after the folds (l > l_split) are on the critical path and small: one element per thread
__host__ __device__ __forceinline__ uint32_t vsum_elems(uint32_t l, uint32_t l_split) { return l <= l_split ? VSUM_ELEMS : VSUM_ELEMS_LATE; }
template <class C> __global__ void __launch_bounds__(TPB_TAIL) k_fold_tail(XyzzPt<C> *F, uint32_t B, uint32_t nb, uint32_t l_first) {
    const uint32_t w = blockIdx.x, i = threadIdx.x;
    for (uint32_t l = l_first; l <= nb; l++) {  // l_first >= 2: the input is always a fold level in F
        const uint32_t m = B >> l;
        if (4 * m <= (uint32_t)TPB_TAIL) {  // few additions left: four lanes per addition
            const uint32_t qd = i >> 2;
            if (qd < m) {
                const XyzzPt<C> *in = F + (size_t)w * B + fold_level_offset(B, l - 1);
                XyzzPt<C> a = in[qd], b = in[qd + m];
                xyzz_add_quad(a, b);
                if ((i & 3) == 0) store_xyzz(&F[(size_t)w * B + fold_level_offset(B, l) + qd], a);
            }
        } else if (i < m) {
            body_fold<C>(w, i, l, B, nullptr, F, nullptr);
        }
        __syncthreads();
    }
}
// Partial sums of the upper halves: grid (chunk, level - l_first, window); a CTA sums up to vsum_elems(l) * TPB_TREE elements of one upper half.
template <class C>
__global__ void __launch_bounds__(TPB_TREE) k_vsum1(const XyzzPt<C> *buckets, const XyzzPt<C> *F, uint32_t B, uint32_t nb, uint32_t chunks_max,
                                                    const uint32_t *bucket_offsets, XyzzPt<C> *partial, uint32_t l_first, uint32_t l_split) {
    __shared__ uint32_t sm[xyzz_words<C>() * TPB_TREE];
    const uint32_t chunk = blockIdx.x, l = blockIdx.y + l_first, w = blockIdx.z;
    const uint32_t m = B >> l, elems = vsum_elems(l, l_split);
    if (chunk * elems * TPB_TREE >= m) return;  // uniform per CTA
    XyzzPt<C> acc = xyzz_identity<C>();
#pragma unroll 1
    for (uint32_t k = 0; k < elems; k++) {
        uint32_t i = chunk * elems * TPB_TREE + k * TPB_TREE + threadIdx.x;
        XyzzPt<C> x;
        if (i < m && fold_upper_elem<C>(w, i, l, B, buckets, F, bucket_offsets, x)) xyzz_add(acc, x);
    }
    acc = block_tree_sum<C>(acc, sm);
    if (threadIdx.x == 0) store_xyzz(&partial[((size_t)w * nb + (l - 1)) * chunks_max + chunk], acc);
}
// grid (level - 1, window): V[w][bit] = sum of that level's partials, bit = nb - level
template <class C>
__global__ void __launch_bounds__(TPB_TREE) k_vsum2(const XyzzPt<C> *partial, uint32_t B, uint32_t nb, uint32_t chunks_max, XyzzPt<C> *V, uint32_t l_split) {
    __shared__ uint32_t sm[xyzz_words<C>() * TPB_TREE];
    const uint32_t l = blockIdx.x + 1, w = blockIdx.y;
    const uint32_t m = B >> l, per = vsum_elems(l, l_split) * TPB_TREE;
    const uint32_t cnt = (m + per - 1) / per;
    const XyzzPt<C> *src = partial + ((size_t)w * nb + (l - 1)) * chunks_max;
    XyzzPt<C> acc = xyzz_identity<C>();
    for (uint32_t i = threadIdx.x; i < cnt; i += TPB_TREE) xyzz_add(acc, src[i]);
    acc = block_tree_sum<C>(acc, sm);
    if (threadIdx.x == 0) store_xyzz(&V[(size_t)w * nb + (nb - l)], acc);
}
// grid (window): out[w] = T0 + sum_b 2^b V[w][b]; quad b doubles V_b b times (four lanes per doubling), then a quad tree sum over the nb + 1 terms.
template <class C> __global__ void __launch_bounds__(TPB_TREE) k_fold_combine(const XyzzPt<C> *F, const XyzzPt<C> *V, uint32_t B, uint32_t nb, XyzzPt<C> *out) {
    __shared__ uint32_t sm[xyzz_words<C>() * TPB_TREE];
    static_assert(TPB_TREE / 4 >= 24, "one quad per term: nb + 1 <= 24");
    const uint32_t w = blockIdx.x, t = threadIdx.x, qd = t >> 2;
    XyzzPt<C> acc = xyzz_identity<C>();
    if (qd == nb) acc = F[(size_t)w * B + (B - 2)];
    else if (qd < nb) {
        acc = V[(size_t)w * nb + qd];
        for (uint32_t d = 0; d < qd; d++) acc = xyzz_dbl_quad(acc);
    }
    if ((t & 3) == 0) sm_put<C>(sm, qd, acc);
    __syncthreads();
    quad_tree_sum<C>(sm, TPB_TREE / 8);
    if (t == 0) store_xyzz(&out[w], sm_get<C>(sm, 0));
}
// Horner over windows on the device (kept for kgr_set_param("final_on_device", 1)); one thread.
template <class C> __global__ void k_final(MsmShape sh, const XyzzPt<C> *win_a, XyzzPt<C> *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    XyzzPt<C> r = win_a[sh.W - 1];
    for (uint32_t w = sh.W - 1; w-- > 0;) {
        for (uint32_t d = 0; d < sh.c; d++) r = xyzz_dbl(r);
        xyzz_add(r, win_a[w]);
    }
    store_xyzz(out, r);
}
template <class C>
__global__ void __launch_bounds__(128) k_precompute(uint32_t n, uint32_t c, uint32_t W, uint32_t stride, const AffinePt<C> *pts, AffinePt<C> *table) {
    body_precompute<C>(blockIdx.x * blockDim.x + threadIdx.x, n, c, W, stride, pts, table);
}
template <class C> __global__ void k_fold_inf(AffinePt<C> *pts, const uint8_t *inf, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !inf[i]) return;
    pts[i].x = El<typename C::Elem>::zero();
    pts[i].y = El<typename C::Elem>::zero();
}

template <class C> __global__ void k_point_op(int op, const AffinePt<C> *a, const AffinePt<C> *b, uint32_t *out24, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XyzzPt<C> acc = xyzz_from_affine(a[i]);
    if (op == 0) xyzz_madd(acc, b[i]);
    else if (op == 1) acc = xyzz_dbl(acc);
    else {
        // go through a non-trivial representative of b: (b + a) - a would need neg; use b doubled path instead
        XyzzPt<C> q = xyzz_from_affine(b[i]);
        XyzzPt<C> t = xyzz_dbl(q);      // 2b
        xyzz_madd(t, b[i]);             // 3b  (non-unit zz)
        xyzz_add(acc, t);               // a + 3b
    }
    typename C::Elem o[3];
    xyzz_to_projective(acc, o);
    constexpr int NW = El<typename C::Elem>::WORDS;
    for (int k = 0; k < 3; k++)
        for (int j = 0; j < NW; j++) out24[3 * NW * (size_t)i + NW * k + j] = El<typename C::Elem>::word(o[k], j);
}

// The quad-cooperative operations against the oracle (four threads per element): op 3: a + 3b with 3b = dbl_quad(b) + b, op 4: (a + b) + (b + a)
// through the equal-points branch on two different representatives, op 5: (a + b) + (-(b + a)) -> identity.
template <class C> __global__ void k_point_op_quad(int op, const AffinePt<C> *a, const AffinePt<C> *b, uint32_t *out24, uint32_t n) {
    uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    if (i >= n) return;
    XyzzPt<C> acc = xyzz_from_affine(a[i]);
    if (op == 3) {
        XyzzPt<C> t = xyzz_dbl_quad(xyzz_from_affine(b[i]));
        xyzz_madd(t, b[i]);
        xyzz_add_quad(acc, t);
    } else {
        xyzz_madd(acc, b[i]);  // a + b, zz = (xb - xa)^2
        XyzzPt<C> v = xyzz_dbl(xyzz_from_affine(b[i]));
        xyzz_madd(v, a[i]);    // 2b + a
        AffinePt<C> nb_ = b[i];
        nb_.y = fp_neg(nb_.y);
        xyzz_madd(v, nb_);     // (2b + a) - b = b + a on another representative (identity operands fall through the special cases)
        if (op == 5) v.y = fp_neg(v.y);
        xyzz_add_quad(acc, v);
    }
    if ((threadIdx.x & 3) != 0) return;
    typename C::Elem o[3];
    xyzz_to_projective(acc, o);
    constexpr int NW = El<typename C::Elem>::WORDS;
    for (int k = 0; k < 3; k++)
        for (int j = 0; j < NW; j++) out24[3 * NW * (size_t)i + NW * k + j] = El<typename C::Elem>::word(o[k], j);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
// k_i = from_u512(8 words of splitmix64(seed, i)) in the curve's scalar field (Montgomery form)
template <class C> __global__ void k_gen_scalars(uint64_t seed, uint64_t first, uint32_t n, Fp<typename C::Scalar> *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t w[16];
    for (int j = 0; j < 8; j++) {
        uint64_t v = splitmix64(seed ^ splitmix64((first + i) * 8 + j));
        w[2 * j] = (uint32_t)v;
        w[2 * j + 1] = (uint32_t)(v >> 32);
    }
    out[i] = fp_from_u512<typename C::Scalar>(w);
}
// Fixed-base scalar multiplication k * G for many k (PedersenCommitment::new, nova/src/pedersen.rs:10-13; CRS setup,
// groth16/src/zksnark.rs:55-57,172-193 — the reference pays one full double-and-add per point).  A table of
// T[j][d] = d * 2^(8 j) * G (32 windows x 256 digits, affine, 512 KB on G1: L2-resident) turns every product into at most 32
// mixed additions and one inversion.
constexpr uint32_t FIXED_WINDOWS = 32, FIXED_DIGITS = 256;
template <class C> KGR_HD AffinePt<C> xyzz_to_affine_fast(const XyzzPt<C> &p) {
    typedef typename C::Elem E;
    AffinePt<C> r;
    if (xyzz_is_identity(p)) {
        r.x = El<E>::zero();
        r.y = El<E>::zero();
        return r;
    }
    E t = fp_inv_fast(fp_mul(p.zz, p.zzz));  // safegcd inversion (modinv.cuh): same unique inverse as Fermat, ~4.5x cheaper
    r.x = fp_mul(p.x, fp_mul(t, p.zzz));
    r.y = fp_mul(p.y, fp_mul(t, p.zz));
    return r;
}
// one thread per table entry (j, d): 8 j doublings of G, then d times that point
template <class C> __global__ void __launch_bounds__(128) k_fixed_table(AffinePt<C> g, AffinePt<C> *table) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= FIXED_WINDOWS * FIXED_DIGITS) return;
    uint32_t j = i / FIXED_DIGITS, d = i % FIXED_DIGITS;
    XyzzPt<C> base = xyzz_from_affine(g);
    for (uint32_t t = 0; t < 8 * j; t++) base = xyzz_dbl(base);
    XyzzPt<C> acc = xyzz_identity<C>();
    for (int bit = 7; bit >= 0; bit--) {
        acc = xyzz_dbl(acc);
        if ((d >> bit) & 1) xyzz_add(acc, base);
    }
    table[i] = xyzz_to_affine_fast(acc);
}
// out[i] = k[i] * G, affine: one table entry per non-zero byte of the canonical scalar
template <class C> __global__ void __launch_bounds__(128) k_fixed_base(const Fp<typename C::Scalar> *k, const AffinePt<C> *table, uint32_t n, AffinePt<C> *out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<typename C::Scalar> s = fp_from_mont(k[i]);
    XyzzPt<C> acc = xyzz_identity<C>();
    for (uint32_t j = 0; j < FIXED_WINDOWS; j++) {
        uint32_t d = (s.v[j >> 2] >> (8 * (j & 3))) & 0xffu;
        if (d) xyzz_madd(acc, load_affine(table, j * FIXED_DIGITS + d));
    }
    out[i] = xyzz_to_affine_fast(acc);
}


}  // namespace kgr
