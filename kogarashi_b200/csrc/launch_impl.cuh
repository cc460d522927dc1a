// launch_impl.cuh — definitions of Launch<C>; include once per curve and instantiate explicitly.
// KGR_PART selects which member groups a translation unit defines (an explicit instantiation of the class instantiates only the
// members defined at that point), so a curve whose kernels compile slowly (G2) can be spread over several TUs:
// 1 scalar side + accumulate + fixup, 2 batched-affine tree levels, 4 running-sum reduce, 8 fold reduce, 16 utilities.
#pragma once
#ifndef KGR_PART
#define KGR_PART 31
#endif
#include "kernels_curve.cuh"
#include "launch.cuh"
#if KGR_PART & 2
#include "affine_kernels.cuh"
#include "scan.cuh"
#endif

namespace kgr {

static inline unsigned cdiv(size_t a, unsigned b) { return (unsigned)((a + b - 1) / b); }

#if KGR_PART & 1
template <class C> int Launch<C>::accumulate_blocks_per_sm() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_accumulate<C>, TPB_ACC, 0);
    return nb;
}
template <class C> void Launch<C>::count(cudaStream_t st, const MsmShape &sh, const uint32_t *scalars, int is_mont, uint32_t *counts, uint32_t *digits) {
    k_count<C><<<cdiv(sh.n, TPB_SCALAR), TPB_SCALAR, 0, st>>>(sh, scalars, is_mont, counts, digits);
}
template <class C>
void Launch<C>::fill_window(cudaStream_t st, const MsmShape &sh, const uint32_t *digits, uint32_t *counts, const uint32_t *offsets, uint32_t *entries) {
    k_fill_window<C><<<dim3(cdiv(sh.n, TPB_SCALAR), sh.W), TPB_SCALAR, 0, st>>>(sh, digits, counts, offsets, entries);
}
template <class C>
void Launch<C>::fill(cudaStream_t st, const MsmShape &sh, const uint32_t *scalars, int is_mont, uint32_t *counts, const uint32_t *offsets, uint32_t *entries) {
    k_fill<C><<<cdiv(sh.n, TPB_SCALAR), TPB_SCALAR, 0, st>>>(sh, scalars, is_mont, counts, offsets, entries);
}
template <class C>
void Launch<C>::accumulate(cudaStream_t st, const MsmShape &sh, uint32_t chunks, const A *bases, const uint32_t *offsets, const uint32_t *entries, X *bucket_acc,
                           X *head, X *tail, uint32_t *tail_bucket) {
    k_accumulate<C><<<cdiv(chunks, TPB_ACC), TPB_ACC, 0, st>>>(sh, bases, offsets, entries, bucket_acc, head, tail, tail_bucket);
}
#endif
#if KGR_PART & 2
static inline uint32_t affine_threads(uint32_t out_max) { return cdiv(cdiv(out_max, AFF_K), AFF_TPB) * AFF_TPB; }
template <class C> void Launch<C>::affine_scratch_words(uint32_t out_max0, size_t &pre_words, size_t &tot_words) {
    const size_t NT = affine_threads(out_max0), EW = El<typename C::Elem>::WORDS;
    pre_words = NT * AFF_K * EW;
    tot_words = 2 * NT * EW;  // totals + the prefix products of k_batch_inv
}
template <class C>
int Launch<C>::affine_levels(cudaStream_t st, uint32_t levels, uint32_t G, const A *bases, const uint32_t *entries, uint32_t *const off[], A *const nodes[2],
                             const uint32_t out_max[], uint32_t *cnt, uint32_t *tile_sums, uint32_t *pre, uint32_t *tot, uint32_t n_bases, void *xs) {
    int launches = 0;
    if (xs) {  // dense copy of the x coordinates for the level-0 denominators (affine_kernels.cuh: AffLevelIn::xs)
        k_extract_x<C><<<cdiv(n_bases, 256), 256, 0, st>>>(bases, n_bases, (typename C::Elem *)xs);
        launches++;
    }
    for (uint32_t l = 0; l < levels; l++) {
        k_affine_counts<<<cdiv((size_t)G + 1, 256), 256, 0, st>>>(off[l], G, cnt);
        exclusive_scan_u32(cnt, off[l + 1], G + 1, tile_sums, st);
        AffLevelIn<C> in{l == 0 ? bases : nodes[(l - 1) & 1], l == 0 ? entries : nullptr, l == 0 ? (const typename C::Elem *)xs : nullptr};
        const uint32_t NT = affine_threads(out_max[l]);
        k_affine_den<C><<<NT / AFF_TPB, AFF_TPB, 0, st>>>(in, off[l], off[l + 1], G, NT, pre, tot);
        k_batch_inv<typename C::Elem><<<cdiv(NT, AFF_TPB * INV_K), AFF_TPB, 0, st>>>(tot, NT, NT, tot + (size_t)NT * El<typename C::Elem>::WORDS, off[l + 1] + G);
        k_affine_add<C><<<NT / AFF_TPB, AFF_TPB, 0, st>>>(in, off[l], off[l + 1], G, NT, pre, tot, nodes[l & 1]);
        launches += 4 + (scan_num_tiles(G + 1) > 1 ? 3 : 1);
    }
    return launches;
}
#endif
#if KGR_PART & 1
// worklist records of the hot buckets: a hot bucket owns at least FIXUP_INLINE_MAX chunks outright and adds one record per FL_SLICE pieces (bound with slack)
template <class C> size_t Launch<C>::fixup_records_max(uint32_t chunks) { return (size_t)chunks / FIXUP_INLINE_MAX + (size_t)chunks / (FL_SLICE / 2) + 4; }
template <class C>
void Launch<C>::fixup(cudaStream_t st, const MsmShape &sh, uint32_t chunks, int sm_count, const uint32_t *offsets, X *bucket_acc, const X *head, const X *tail,
                      const uint32_t *tail_bucket, uint32_t *worklist, uint32_t *worklist_len, X *partial, uint32_t *counter) {
    // four lanes per chunk, or per bucket when the chunks outnumber the buckets (kernels_curve.cuh)
    if (chunks > sh.G) k_fixup_buckets<C><<<cdiv((size_t)sh.G * 4, TPB_RED), TPB_RED, 0, st>>>(sh, offsets, bucket_acc, head, tail, worklist, worklist_len);
    else k_fixup<C><<<cdiv((size_t)chunks * 4, TPB_RED), TPB_RED, 0, st>>>(sh, offsets, bucket_acc, head, tail, tail_bucket, worklist, worklist_len);
    unsigned blocks = sh.G < 4u * (unsigned)sm_count ? sh.G : 4u * (unsigned)sm_count;
    if (blocks < (unsigned)FL_MAX_SLICES) blocks = FL_MAX_SLICES;  // a single hot bucket still spreads over its slices
    k_fixup_long<C><<<blocks, TPB_TREE, 0, st>>>(sh, offsets, bucket_acc, head, tail, worklist, worklist_len, partial, counter);
}
#endif
#if KGR_PART & 1
template <class C> void Launch<C>::bucket_merge(cudaStream_t st, uint32_t G, const uint32_t *piece_offsets, const X *piece_acc, X *bucket_acc) {
    k_bucket_merge<C><<<cdiv(G, TPB_RED), TPB_RED, 0, st>>>(G, piece_offsets, piece_acc, bucket_acc);
}
#endif
#if KGR_PART & 4
template <class C>
void Launch<C>::reduce(cudaStream_t st, uint32_t n_windows, uint32_t cnt_in, uint32_t K, uint32_t m_log2, const X *in_s, const X *in_a, X *out_s, X *out_a,
                       const uint32_t *bucket_offsets) {
    uint32_t threads = n_windows * ((cnt_in + K - 1) / K);
    k_reduce<C><<<cdiv(threads, TPB_RED), TPB_RED, 0, st>>>(n_windows, cnt_in, K, m_log2, in_s, in_a, out_s, out_a, bucket_offsets);
}
template <class C> void Launch<C>::weight(cudaStream_t st, uint32_t n_windows, uint32_t cnt, uint32_t m_log2, const X *in_s, const X *in_a, X *out) {
    k_weight<C><<<cdiv((size_t)n_windows * cnt, TPB_RED), TPB_RED, 0, st>>>(n_windows, cnt, m_log2, in_s, in_a, out);
}
template <class C> void Launch<C>::tree_sum(cudaStream_t st, uint32_t n_windows, const X *in, uint32_t cnt_in, X *out) {
    k_tree_sum<C><<<dim3(cdiv(cnt_in, TPB_TREE), n_windows), TPB_TREE, 0, st>>>(in, cnt_in, out);
}
#endif
#if KGR_PART & 8
// columns of `partial` per (window, level): the early levels (<= l_split) are summed VSUM_ELEMS * TPB_TREE elements per CTA, the late ones
// VSUM_ELEMS_LATE * TPB_TREE; level 1 (B / 2 elements) bounds both
template <class C> uint32_t Launch<C>::fold_chunks_max(uint32_t B) { return ((B >> 1) + VSUM_ELEMS_LATE * TPB_TREE - 1) / (VSUM_ELEMS_LATE * TPB_TREE); }
template <class C>
int Launch<C>::fold_reduce(cudaStream_t st, cudaStream_t st2, cudaEvent_t ev_fork, cudaEvent_t ev_join, uint32_t n_windows, uint32_t B, const X *buckets,
                           const uint32_t *bucket_offsets, X *F, X *partial, X *V, X *out) {
    uint32_t nb = 0;
    while ((1u << nb) < B) nb++;
    int launches = 0;
    const uint32_t chunks_max = fold_chunks_max(B);
    const uint32_t l_early = 3;  // upper halves of levels 1..3 are summed on st2 as soon as fold level 2 exists
    const bool fork = l_early < nb && (B >> 2) > (uint32_t)TPB_TAIL;  // fold level 2 gets its own launch
    const uint32_t l_split = fork ? l_early : 0;
    uint32_t l = 1;
    for (; l <= nb; l++) {
        uint32_t m = B >> l;
        if (l >= 2 && m <= (uint32_t)TPB_TAIL) break;  // the rest in one kernel
        // four lanes per addition once 4 x the additions fit the resident threads of the GPU (~75 K): 2^20 points, levels 4 and 5
        if (l >= 2 && (size_t)n_windows * m <= 18000) k_fold_quad<C><<<cdiv((size_t)n_windows * m * 4, TPB_RED), TPB_RED, 0, st>>>(F, B, l, n_windows);
        else k_fold<C><<<cdiv((size_t)n_windows * m, TPB_RED), TPB_RED, 0, st>>>(buckets, F, B, l, n_windows, bucket_offsets);
        launches++;
        if (fork && l + 1 == l_early) {
            cudaEventRecord(ev_fork, st);
            cudaStreamWaitEvent(st2, ev_fork, 0);
            const uint32_t chunks = ((B >> 1) + VSUM_ELEMS * TPB_TREE - 1) / (VSUM_ELEMS * TPB_TREE);
            k_vsum1<C><<<dim3(chunks, l_early, n_windows), TPB_TREE, 0, st2>>>(buckets, F, B, nb, chunks_max, bucket_offsets, partial, 1, l_split);
            cudaEventRecord(ev_join, st2);
            launches++;
        }
    }
    if (l <= nb) {
        k_fold_tail<C><<<n_windows, TPB_TAIL, 0, st>>>(F, B, nb, l);
        launches++;
    }
    {
        uint32_t first = fork ? l_early + 1 : 1;
        uint32_t m = B >> first;
        uint32_t chunks = (m + VSUM_ELEMS_LATE * TPB_TREE - 1) / (VSUM_ELEMS_LATE * TPB_TREE);
        k_vsum1<C><<<dim3(chunks, nb - first + 1, n_windows), TPB_TREE, 0, st>>>(buckets, F, B, nb, chunks_max, bucket_offsets, partial, first, l_split);
        launches++;
    }
    if (fork) cudaStreamWaitEvent(st, ev_join, 0);
    k_vsum2<C><<<dim3(nb, n_windows), TPB_TREE, 0, st>>>(partial, B, nb, chunks_max, V, l_split);
    k_fold_combine<C><<<n_windows, TPB_TREE, 0, st>>>(F, V, B, nb, out);
    return launches + 2;
}
#endif
#if KGR_PART & 4
template <class C> void Launch<C>::final_horner(cudaStream_t st, const MsmShape &sh, const X *win_a, X *out) { k_final<C><<<1, 32, 0, st>>>(sh, win_a, out); }
#endif
#if KGR_PART & 16
template <class C> void Launch<C>::fold_inf(cudaStream_t st, A *pts, const uint8_t *inf, uint32_t n) { k_fold_inf<C><<<cdiv(n, 256), 256, 0, st>>>(pts, inf, n); }
template <class C> void Launch<C>::precompute(cudaStream_t st, uint32_t n, uint32_t c, uint32_t W, uint32_t stride, const A *pts, A *table) {
    k_precompute<C><<<cdiv(n, 128), 128, 0, st>>>(n, c, W, stride, pts, table);
}
template <class C> void Launch<C>::point_op(cudaStream_t st, int op, const A *a, const A *b, uint32_t *out24, uint32_t n) {
    if (op >= 3) k_point_op_quad<C><<<cdiv((size_t)n * 4, 64), 64, 0, st>>>(op, a, b, out24, n);
    else k_point_op<C><<<cdiv(n, 64), 64, 0, st>>>(op, a, b, out24, n);
}
template <class C> void Launch<C>::gen_scalars(cudaStream_t st, uint64_t seed, uint64_t first, uint32_t n, S *out) {
    k_gen_scalars<C><<<cdiv(n, 256), 256, 0, st>>>(seed, first, n, out);
}
template <class C> size_t Launch<C>::fixed_table_points() { return (size_t)FIXED_WINDOWS * FIXED_DIGITS; }
template <class C> void Launch<C>::fixed_table(cudaStream_t st, const A &g, A *table) { k_fixed_table<C><<<cdiv(FIXED_WINDOWS * FIXED_DIGITS, 128), 128, 0, st>>>(g, table); }
template <class C> void Launch<C>::fixed_base(cudaStream_t st, const S *k, const A *table, uint32_t n, A *out) { k_fixed_base<C><<<cdiv(n, 128), 128, 0, st>>>(k, table, n, out); }
#endif

}  // namespace kgr
