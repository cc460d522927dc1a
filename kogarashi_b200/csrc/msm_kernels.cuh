// msm_kernels.cuh — per-thread bodies of the MSM pipeline.
//
// Replaces the hot loops of groth16/src/msm.rs:6-48 (bucket fill :25-30, running sum :35-40,
// window shift :41, fold :45-47) with a bucket-sorted pipeline:
//
//   recode   canonical scalar -> W signed c-bit digits (|d| <= 2^(c-1)), one bucket id per digit
//   count    histogram of bucket ids            (global bucket g = window * B + |d| - 1)
//   scan     exclusive prefix sum -> offsets[g]  (scan.cuh)
//   fill     counting-sort scatter of (point index | sign << 31) into entries[]
//   accumulate  thread t owns entries [t*L, (t+1)*L): XYZZ mixed adds into a register accumulator,
//            flushing at bucket boundaries; buckets cut by a chunk boundary go to head/tail slots
//   fixup    stitches the head/tail partial sums of buckets that span chunks
//   reduce   hierarchical running sums: groups of K, (S, A) = (sum, weighted sum), log_K(B) levels
//   final    Horner over windows with c doublings per step, XYZZ -> projective
//
// The reference cuts unsigned digits from the canonical scalar (msm.rs:26,75-91) and uses
// c = f(len) (:7-14); signed digits / a different c change only the representative, never the
// group element.  Every body takes an explicit thread id so tests/host_emu.cpp can execute the
// identical logic on the CPU.
#pragma once
#include "curve.cuh"
#include "modinv.cuh"

namespace kgr {

struct MsmShape {
    uint32_t n;  // scalar-point pairs (min(len) zip semantics are applied by the host layer)
    uint32_t c;  // window bits
    uint32_t W;  // windows = ceil(255 / c)  (scalars < 2^254, one spare bit for the signed carry)
    uint32_t B;  // buckets per window = 2^(c-1)
    uint32_t G;  // W * B
    uint32_t L;  // entries per accumulate thread (at most: see eff_chunk_len)
    uint32_t K;  // reduce fan-in
    // Window-collapsed mode (bases registered with kgr_bases_precompute): the table holds 2^(c*w) * P_i for
    // every window w at index w * pstride + i, so all windows share ONE set of B buckets (gstride = 0,
    // G = B) and no doublings remain after the bucket reduction.  Normal mode: gstride = B, pstride = 0.
    uint32_t gstride;  // global bucket id = w * gstride + bucket
    uint32_t pstride;  // entry payload  = w * pstride + poff + i
    uint32_t poff;
    // Threads of the accumulate grid the host launched for its BOUND on the entries (n * W, or the node bound after the batched-affine levels).
    // 0: every chunk is exactly L entries.  Otherwise the chunk length follows the number of entries actually on the device (eff_chunk_len).
    uint32_t chunks = 0;
};

KGR_HD uint32_t window_raw(const uint32_t s[8], uint32_t bit, uint32_t c) {
    uint32_t limb = bit >> 5, sh = bit & 31;
    if (limb >= 8) return 0;
    uint64_t v = s[limb];
    if (limb + 1 < 8) v |= (uint64_t)s[limb + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

// if (x >= p) x -= p, five times: any 256-bit input lands in [0, p) (2^256 < 6p for both moduli).  The reference never holds
// an unreduced field element; a caller that passes one (canonical bytes >= the modulus, or limbs that are not fully reduced) gets
// the scalar modulo the modulus instead of a digit string that silently lost its top bits.
template <class P> KGR_HD void reduce_range(Fp<P> &s) {
    for (int k = 0; k < 5; k++) fp_final_sub<P>(s.v);
}

// Load scalar i as canonical limbs.  Montgomery inputs (the reference's in-memory form, fr.rs:71)
// are reduced exactly like to_raw_bytes does (zkstd/src/macros/field.rs:102-104 -> fr.rs:74-84).
template <class C> KGR_HD void load_scalar(const uint32_t *scalars, uint32_t i, int is_mont, uint32_t out[8]) {
    Fp<typename C::Scalar> s;
#if defined(__CUDA_ARCH__)
    const uint4 *p = reinterpret_cast<const uint4 *>(scalars) + 2 * (size_t)i;
    uint4 lo = p[0], hi = p[1];
    s.v[0] = lo.x; s.v[1] = lo.y; s.v[2] = lo.z; s.v[3] = lo.w;
    s.v[4] = hi.x; s.v[5] = hi.y; s.v[6] = hi.z; s.v[7] = hi.w;
#else
    for (int k = 0; k < 8; k++) s.v[k] = scalars[8 * (size_t)i + k];
#endif
    reduce_range(s);
    if (is_mont) s = fp_from_mont(s);
    for (int k = 0; k < 8; k++) out[k] = s.v[k];
}

// Signed recoding.  Calls f(window, bucket_in_window, sign) for each non-zero digit of the windows
// below the top one and returns the top window's digit as (bucket | sign << 31), NO_DIGIT if zero.
// The top window is returned separately because it is structurally hot: it only holds the
// 254 - c*(W-1) leading scalar bits, so a handful of buckets receive n / 2^t entries each.
constexpr uint32_t NO_DIGIT = 0xffffffffu;
template <class Fn> KGR_HD uint32_t for_each_digit(const uint32_t s[8], const MsmShape &sh, Fn f) {
    uint32_t carry = 0, top = NO_DIGIT;
    for (uint32_t w = 0; w < sh.W; w++) {
        uint32_t d = window_raw(s, w * sh.c, sh.c) + carry;
        carry = 0;
        uint32_t sign = 0;
        if (d > sh.B) {
            d = (1u << sh.c) - d;
            sign = 1;
            carry = 1;
        }
        if (d != 0) {
            if (w + 1 == sh.W) top = (d - 1) | (sign << 31);
            else f(w, d - 1, sign);
        }
    }
    return top;
}

KGR_HD uint32_t atomic_add_u32(uint32_t *p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    uint32_t o = *p;
    *p = o + v;
    return o;
#endif
}
KGR_HD uint32_t atomic_sub_u32(uint32_t *p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicSub(p, v);
#else
    uint32_t o = *p;
    *p = o - v;
    return o;
#endif
}

// Warp-aggregated "take `1` from counter[key]" for keys that collide inside a warp: lanes with the
// same key elect a leader (match.any), the leader does one atomic for the whole group and every
// lane derives its own slot from its rank in the group.  `sub`: counts are consumed (fill) instead of
// incremented (count).  Returns the value the counter had for this lane's unit (fill uses it).
// Must be called by all 32 lanes of the warp; lanes without work pass key = NO_DIGIT.
KGR_HD uint32_t warp_aggregated_take(uint32_t *counters, uint32_t key, bool sub) {
#if defined(__CUDA_ARCH__)
    unsigned group = __match_any_sync(0xffffffffu, key);
    if (key == NO_DIGIT) return 0;
    unsigned lane = threadIdx.x & 31;
    unsigned leader = __ffs(group) - 1;
    unsigned rank = __popc(group & ((1u << lane) - 1u));
    unsigned cnt = __popc(group);
    uint32_t base = 0;
    if (lane == leader) base = sub ? atomicSub(&counters[key], cnt) - cnt : atomicAdd(&counters[key], cnt);
    base = __shfl_sync(group, base, leader);
    return base + rank;
#else
    if (key == NO_DIGIT) return 0;
    return sub ? atomic_sub_u32(&counters[key], 1u) - 1u : atomic_add_u32(&counters[key], 1u);
#endif
}

// ---- count / fill ---------------------------------------------------------------------------
// i may be >= n (whole warps run so that the warp-aggregated top window sees all 32 lanes).
// digits != nullptr: also store every window's digit word (bucket | sign << 31, NO_DIGIT for zero) window-major at
// digits[w * n + i], so that the fill pass can run window by window without recoding (body_fill_window).
template <class C>
KGR_HD void body_count(uint32_t i, const MsmShape &sh, const uint32_t *scalars, int is_mont, uint32_t *counts, uint32_t *digits) {
    uint32_t top = NO_DIGIT;
    if (i < sh.n) {
        uint32_t s[8];
        load_scalar<C>(scalars, i, is_mont, s);
        if (digits) {
            uint32_t carry = 0;
            for (uint32_t w = 0; w < sh.W; w++) {
                uint32_t d = window_raw(s, w * sh.c, sh.c) + carry;
                carry = 0;
                uint32_t sign = 0;
                if (d > sh.B) {
                    d = (1u << sh.c) - d;
                    sign = 1;
                    carry = 1;
                }
                uint32_t word = d ? ((d - 1) | (sign << 31)) : NO_DIGIT;
                digits[(size_t)w * sh.n + i] = word;
                if (d) {
                    if (w + 1 == sh.W) top = word;
                    else atomic_add_u32(&counts[w * sh.gstride + d - 1], 1u);
                }
            }
        } else {
            top = for_each_digit(s, sh, [&](uint32_t w, uint32_t b, uint32_t) { atomic_add_u32(&counts[w * sh.gstride + b], 1u); });
        }
    }
    uint32_t key = top == NO_DIGIT ? NO_DIGIT : (sh.W - 1) * sh.gstride + (top & 0x7fffffffu);
    (void)warp_aggregated_take(counts, key, false);
}

// counts[] is consumed back to zero (positions are handed out from the end of each bucket), so
// the histogram buffer is clean for the next MSM without a memset.  Windows are processed eight at
// a time: eight digits, then eight independent atomics, then eight independent stores, so that the
// L2 round trips of one scalar overlap instead of forming a chain.
template <class C>
KGR_HD void body_fill(uint32_t i, const MsmShape &sh, const uint32_t *scalars, int is_mont, uint32_t *counts, const uint32_t *offsets,
                      uint32_t *entries) {
    uint32_t top = NO_DIGIT;
    if (i < sh.n) {
        uint32_t s[8];
        load_scalar<C>(scalars, i, is_mont, s);
        uint32_t carry = 0;
        for (uint32_t w0 = 0; w0 < sh.W; w0 += 8) {
            uint32_t g[8], pay[8], pos[8];
#pragma unroll
            for (uint32_t j = 0; j < 8; j++) {
                uint32_t w = w0 + j;
                g[j] = NO_DIGIT;
                pay[j] = 0;
                if (w < sh.W) {
                    uint32_t d = window_raw(s, w * sh.c, sh.c) + carry;
                    carry = 0;
                    uint32_t sign = 0;
                    if (d > sh.B) {
                        d = (1u << sh.c) - d;
                        sign = 1;
                        carry = 1;
                    }
                    if (d != 0) {
                        if (w + 1 == sh.W) top = (d - 1) | (sign << 31);
                        else {
                            g[j] = w * sh.gstride + d - 1;
                            pay[j] = (w * sh.pstride + sh.poff + i) | (sign << 31);
                        }
                    }
                }
            }
#pragma unroll
            for (uint32_t j = 0; j < 8; j++)
                if (g[j] != NO_DIGIT) pos[j] = offsets[g[j]] + atomic_sub_u32(&counts[g[j]], 1u) - 1u;
#pragma unroll
            for (uint32_t j = 0; j < 8; j++)
                if (g[j] != NO_DIGIT) entries[pos[j]] = pay[j];
        }
    }
    uint32_t key = top == NO_DIGIT ? NO_DIGIT : (sh.W - 1) * sh.gstride + (top & 0x7fffffffu);
    uint32_t k = warp_aggregated_take(counts, key, true);
    if (key != NO_DIGIT) entries[offsets[key] + k] = ((sh.W - 1) * sh.pstride + sh.poff + i) | (top & 0x80000000u);
}

// Window-major fill: thread (i, w) places the entry of scalar i for window w.  All CTAs that run at the same
// time work on the same one or two windows, so the scattered 4-byte stores land in a region of n * 4 bytes
// (L2 resident up to n ~ 2^24) and are merged there instead of costing a DRAM sector each.
template <class C>
KGR_HD void body_fill_window(uint32_t i, uint32_t w, const MsmShape &sh, const uint32_t *digits, uint32_t *counts, const uint32_t *offsets,
                             uint32_t *entries) {
    uint32_t d = (i < sh.n) ? digits[(size_t)w * sh.n + i] : NO_DIGIT;
    uint32_t key = d == NO_DIGIT ? NO_DIGIT : w * sh.gstride + (d & 0x7fffffffu);
    uint32_t pay = (w * sh.pstride + sh.poff + i) | (d & 0x80000000u);
    if (w + 1 == sh.W) {  // structurally hot window: warp-aggregated (all 32 lanes call it)
        uint32_t k = warp_aggregated_take(counts, key, true);
        if (key != NO_DIGIT) entries[offsets[key] + k] = pay;
    } else if (key != NO_DIGIT) {
        entries[offsets[key] + atomic_sub_u32(&counts[key], 1u) - 1u] = pay;
    }
}

// ---- accumulate -----------------------------------------------------------------------------
template <class C> KGR_HD AffinePt<C> load_affine(const AffinePt<C> *bases, uint32_t idx) {
    AffinePt<C> p;
    el_load(p.x, &bases[idx].x);
    el_load(p.y, &bases[idx].y);
    return p;
}

template <class C> KGR_HD void store_xyzz(XyzzPt<C> *dst, const XyzzPt<C> &p) {
    el_store(&dst->x, p.x);
    el_store(&dst->y, p.y);
    el_store(&dst->zz, p.zz);
    el_store(&dst->zzz, p.zzz);
}

// Chunk length actually used.  With fewer entries on the device than the host's bound (zero digits, and above all skewed scalars: Nova witnesses
// are mostly 0 / 1) the same grid takes shorter chunks, down to 4: the chain of L dependent additions per thread is the latency of the
// accumulate kernel (2^20 points, all scalars 0 / 1: 131 K nodes left for 70 K threads that would otherwise run 2 K chunks of 64).
KGR_HD uint32_t eff_chunk_len(const MsmShape &sh, uint32_t M) {
    if (!sh.chunks) return sh.L;
    uint32_t L = (uint32_t)(((uint64_t)M + sh.chunks - 1) / sh.chunks);
    return L < 4 ? 4 : (L > sh.L ? sh.L : L);
}

// first g with offsets[g+1] > pos  (offsets is non-decreasing, offsets[G] = M > pos)
KGR_HD uint32_t bucket_of_position(const uint32_t *offsets, uint32_t G, uint32_t pos) {
    uint32_t lo = 0, hi = G;  // answer in [lo, hi)
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (offsets[mid] <= pos) lo = mid;
        else hi = mid;
    }
    return lo;
}

// A segment that started before this chunk goes to head[t]; one that started here and continues past
// the chunk goes to tail[t] (at most one per thread: the last segment) and its bucket id is returned
// so that the fix-up pass can start from the chunks instead of scanning every bucket.
template <class C>
KGR_HD uint32_t flush_segment(uint32_t t, uint32_t g, uint32_t s, uint32_t e, const uint32_t *offsets, const XyzzPt<C> &acc, XyzzPt<C> *bucket_acc,
                              XyzzPt<C> *head, XyzzPt<C> *tail) {
    uint32_t lo = offsets[g], hi = offsets[g + 1];
    if (lo < s) store_xyzz(&head[t], acc);
    else if (hi > e) {
        store_xyzz(&tail[t], acc);
        return g;
    } else store_xyzz(&bucket_acc[g], acc);
    return NO_DIGIT;
}

template <class C>
KGR_HD void body_accumulate(uint32_t t, const MsmShape &sh, const AffinePt<C> *bases, const uint32_t *offsets, const uint32_t *entries,
                            XyzzPt<C> *bucket_acc, XyzzPt<C> *head, XyzzPt<C> *tail, uint32_t *tail_bucket) {
    uint32_t M = offsets[sh.G];
    const uint32_t L = eff_chunk_len(sh, M);
    uint64_t s64 = (uint64_t)t * L;
    if (s64 >= M) return;
    uint32_t s = (uint32_t)s64;
    uint32_t e = (M - s > L) ? s + L : M;
    uint32_t g = bucket_of_position(offsets, sh.G, s);
    uint32_t g_end = offsets[g + 1];
    XyzzPt<C> acc = xyzz_identity<C>();
    // entries == nullptr: the "bases" are the nodes left by the batched-affine levels (affine_kernels.cuh), one per position, no sign
    uint32_t ent = entries ? entries[s] : s;
    AffinePt<C> pt = load_affine(bases, ent & 0x7fffffffu);
    for (uint32_t pos = s; pos < e; pos++) {
        // prefetch the next entry's point while this one is being added
        uint32_t ent_next = 0;
        AffinePt<C> pt_next = pt;
        if (pos + 1 < e) {
            ent_next = entries ? entries[pos + 1] : pos + 1;
            pt_next = load_affine(bases, ent_next & 0x7fffffffu);
        }
        if (pos >= g_end) {
            (void)flush_segment(t, g, s, e, offsets, acc, bucket_acc, head, tail);  // ends inside the chunk: never a tail
            acc = xyzz_identity<C>();
            do {
                g++;
                g_end = offsets[g + 1];
            } while (pos >= g_end);
        }
        pt.y = fp_cneg(pt.y, (ent >> 31) != 0);
        xyzz_madd(acc, pt);
        ent = ent_next;
        pt = pt_next;
    }
    tail_bucket[t] = flush_segment(t, g, s, e, offsets, acc, bucket_acc, head, tail);
}

// ---- helpers of the batched-affine tree levels (affine_kernels.cuh) -------------------------------------------------------------------------
// x coordinate only (32 bytes): all the chord denominators of phase 1 need
template <class C> KGR_HD typename C::Elem load_x(const AffinePt<C> *p) {
    typename C::Elem x;
    el_load(x, &p->x);
    return x;
}
// case of the pair (a, b) and its denominator: 0 chord (xb - xa), 1 tangent (2 ya), 2 result a, 3 result b, 4 result identity
template <class C> KGR_HD int pair_case(const AffinePt<C> &a, const AffinePt<C> &b, typename C::Elem &den) {
    typedef typename C::Elem E;
    den = El<E>::one();
    if (affine_is_identity(b)) return 2;
    if (affine_is_identity(a)) return 3;
    E dx = fp_sub(b.x, a.x);
    if (!fp_is_zero(dx)) {
        den = dx;
        return 0;
    }
    if (fp_eq(a.y, b.y) && !fp_is_zero(a.y)) {
        den = fp_dbl(a.y);
        return 1;
    }
    return 4;
}
template <class C> KGR_HD AffinePt<C> pair_sum(int code, const AffinePt<C> &a, const AffinePt<C> &b, const typename C::Elem &den_inv) {
    typedef typename C::Elem E;
    if (code == 2) return a;
    if (code == 3) return b;
    AffinePt<C> r;
    if (code == 4) {
        r.x = El<E>::zero();
        r.y = El<E>::zero();
        return r;
    }
    E num;
    if (code == 0) num = fp_sub(b.y, a.y);
    else {
        E xx = fp_sqr(a.x);
        num = fp_add(fp_dbl(xx), xx);
    }
    E lam = fp_mul(num, den_inv);
    r.x = fp_sub(fp_sub(fp_sqr(lam), a.x), b.x);
    r.y = fp_sub(fp_mul(lam, fp_sub(a.x, r.x)), a.y);
    return r;
}

// One thread per chunk: a chunk whose last segment continues into the following chunks owns that bucket
// and sums its pieces (its own tail slot, then the head slots of the following chunks).  Buckets cut
// into more than FIXUP_INLINE_MAX pieces (a hot bucket: skewed scalars, or a thin top window) are
// queued for the block-cooperative kernel instead of being walked by one thread.  Empty buckets are
// never written: the first reduce level recognises them from offsets[].
constexpr uint32_t FIXUP_INLINE_MAX = 6;
template <class C>
KGR_HD void body_fixup(uint32_t t, const MsmShape &sh, const uint32_t *offsets, XyzzPt<C> *bucket_acc, const XyzzPt<C> *head,
                       const XyzzPt<C> *tail, const uint32_t *tail_bucket, uint32_t *worklist, uint32_t *worklist_len) {
    uint32_t M = offsets[sh.G];
    const uint32_t L = eff_chunk_len(sh, M);
    if ((uint64_t)t * L >= M) return;
    uint32_t g = tail_bucket[t];
    if (g == NO_DIGIT) return;
    uint32_t t1 = (offsets[g + 1] - 1) / L;
    if (t1 - t > FIXUP_INLINE_MAX && worklist) {
        worklist[atomic_add_u32(worklist_len, 1u)] = g;
        return;
    }
    XyzzPt<C> acc = tail[t];
    for (uint32_t u = t + 1; u <= t1; u++) xyzz_add(acc, head[u]);
    store_xyzz(&bucket_acc[g], acc);
}

// A host-buffer call streamed in pieces (msm.cu: enqueue_msm, StreamPiece): every piece is accumulated into its own bucket array and then
// added, bucket by bucket, to the call's zero-initialised bucket set (all-zero words are an XYZZ identity: zz = 0), so that ONE bucket
// reduction serves the whole call.  One thread per bucket: every lane does the same general addition, unlike an addition at the moment a
// bucket is flushed inside the accumulate loop, which diverges (measured: the accumulate kernel 2.3x slower, profiles/r02_e2e.md).
template <class C> KGR_HD void body_bucket_merge(uint32_t g, uint32_t G, const uint32_t *piece_offsets, const XyzzPt<C> *piece_acc, XyzzPt<C> *bucket_acc) {
    if (g >= G || piece_offsets[g] == piece_offsets[g + 1]) return;  // buckets the piece never wrote
    XyzzPt<C> a = bucket_acc[g];
    xyzz_add(a, piece_acc[g]);
    store_xyzz(&bucket_acc[g], a);
}

// Partial sum of the pieces of hot bucket g owned by lane `lane` of `lanes` cooperating threads
// (piece 0 is tail[t0], piece k >= 1 is head[t0 + k]); the caller tree-adds the lane results.
template <class C>
KGR_HD XyzzPt<C> fixup_long_partial(uint32_t g, uint32_t lane, uint32_t lanes, const MsmShape &sh, const uint32_t *offsets, const XyzzPt<C> *head,
                                    const XyzzPt<C> *tail) {
    uint32_t lo = offsets[g], hi = offsets[g + 1];
    const uint32_t L = eff_chunk_len(sh, offsets[sh.G]);
    uint32_t t0 = lo / L, t1 = (hi - 1) / L;
    XyzzPt<C> acc = xyzz_identity<C>();
    for (uint32_t t = t0 + lane; t <= t1; t += lanes) xyzz_add(acc, t == t0 ? tail[t0] : head[t]);
    return acc;
}

// Table for the window-collapsed mode: table[w * stride + i] = 2^(c*w) * P_i as an affine point.
template <class C> KGR_HD void body_precompute(uint32_t i, uint32_t n, uint32_t c, uint32_t W, uint32_t stride, const AffinePt<C> *pts, AffinePt<C> *table) {
    if (i >= n) return;
    XyzzPt<C> acc = xyzz_from_affine(pts[i]);
    for (uint32_t w = 0; w < W; w++) {
        table[(size_t)w * stride + i] = xyzz_to_affine(acc);
        if (w + 1 < W)
            for (uint32_t d = 0; d < c; d++) acc = xyzz_dbl(acc);
    }
}

// ---- reduce ---------------------------------------------------------------------------------
// Level l combines K consecutive elements of a window, each standing for m = 2^m_log2 buckets
// with (s_i, a_i) = (plain sum, sum weighted 1..m):
//   S = sum s_i ;  A = sum a_i + m * sum_i i*s_i          (i = 0..K-1)
// Level 0 has a_i == s_i == bucket i (in_a == nullptr).
template <class C>
KGR_HD void body_reduce(uint32_t tid, uint32_t n_windows, uint32_t cnt_in, uint32_t K, uint32_t m_log2, const XyzzPt<C> *in_s,
                        const XyzzPt<C> *in_a, XyzzPt<C> *out_s, XyzzPt<C> *out_a, const uint32_t *bucket_offsets) {
    uint32_t cnt_out = (cnt_in + K - 1) / K;
    if (tid >= n_windows * cnt_out) return;
    uint32_t w = tid / cnt_out, gi = tid % cnt_out;
    uint32_t base = gi * K;
    uint32_t k = (cnt_in - base < K) ? cnt_in - base : K;
    const XyzzPt<C> *s_in = in_s + (size_t)w * cnt_in + base;
    // level 0 only: bucket_offsets != nullptr, an empty bucket (never written by accumulate / fixup) is the identity
    const uint32_t *off = bucket_offsets ? bucket_offsets + (size_t)w * cnt_in + base : nullptr;
    XyzzPt<C> run = xyzz_identity<C>(), T = xyzz_identity<C>();
    for (uint32_t i = k; i-- > 1;) {
        if (!off || off[i] != off[i + 1]) xyzz_add(run, s_in[i]);
        xyzz_add(T, run);
    }
    if (!off || off[0] != off[1]) xyzz_add(run, s_in[0]);  // run = S
    XyzzPt<C> A;
    if (in_a) {
        const XyzzPt<C> *a_in = in_a + (size_t)w * cnt_in + base;
        A = a_in[0];
        for (uint32_t i = 1; i < k; i++) xyzz_add(A, a_in[i]);
    } else {
        A = run;
    }
    for (uint32_t d = 0; d < m_log2; d++) T = xyzz_dbl(T);
    xyzz_add(A, T);
    store_xyzz(&out_s[(size_t)w * cnt_out + gi], run);
    store_xyzz(&out_a[(size_t)w * cnt_out + gi], A);
}

// Closes the reduction once few elements per window are left: element g of a window stands for
// m = 2^m_log2 buckets starting at bucket g*m, with (s_g, a_g) = (plain sum, sum weighted 1..m), so
// its contribution to the window sum is a_g + m*g*s_g; g*s_g by MSB-first double-and-add.  All
// elements are independent, the results are then tree-summed per window (k_tree_sum).
template <class C>
KGR_HD void body_weight(uint32_t tid, uint32_t n_windows, uint32_t cnt, uint32_t m_log2, const XyzzPt<C> *in_s, const XyzzPt<C> *in_a,
                        XyzzPt<C> *out) {
    if (tid >= n_windows * cnt) return;
    uint32_t g = tid % cnt;
    XyzzPt<C> s = in_s[tid];
    XyzzPt<C> acc = xyzz_identity<C>();
    if (g != 0) {
        int top = 31;
        while (!((g >> top) & 1)) top--;
        for (int bit = top; bit >= 0; bit--) {
            acc = xyzz_dbl(acc);
            if ((g >> bit) & 1) xyzz_add(acc, s);
        }
        for (uint32_t d = 0; d < m_log2; d++) acc = xyzz_dbl(acc);
    }
    xyzz_add(acc, in_a ? in_a[tid] : s);
    store_xyzz(&out[tid], acc);
}

// ---- fold reduce: index arithmetic shared by the kernels (kernels_curve.cuh) and the host emulation (tests/host_emu.cpp) --------------------
// Row of one window in F: level l (m = B >> l elements) lives at offset B - 2m = B - (B >> (l - 1)); level 0 is the bucket array itself.
KGR_HD uint32_t fold_level_offset(uint32_t B, uint32_t l) { return B - (B >> (l - 1)); }
// One element of fold level l: Y^(l)[i] = Y^(l-1)[i] + Y^(l-1)[i + m].  Level 1 reads the buckets and skips the ones the accumulate kernel never
// wrote (equal offsets).
template <class C>
KGR_HD void body_fold(uint32_t w, uint32_t i, uint32_t l, uint32_t B, const XyzzPt<C> *buckets, XyzzPt<C> *F, const uint32_t *bucket_offsets) {
    const uint32_t m = B >> l;
    const XyzzPt<C> *in = (l == 1) ? buckets + (size_t)w * B : F + (size_t)w * B + fold_level_offset(B, l - 1);
    bool lo_ok = true, hi_ok = true;
    if (l == 1 && bucket_offsets) {
        const uint32_t *off = bucket_offsets + (size_t)w * B;
        lo_ok = off[i] != off[i + 1];
        hi_ok = off[i + m] != off[i + m + 1];
    }
    XyzzPt<C> a = lo_ok ? in[i] : xyzz_identity<C>();
    if (hi_ok) {
        XyzzPt<C> b = in[i + m];
        xyzz_add(a, b);
    }
    store_xyzz(&F[(size_t)w * B + fold_level_offset(B, l) + i], a);
}
// Element i of the upper half that level l folds away (its plain sum is V_{nb - l}); false if it is an empty bucket.
template <class C>
KGR_HD bool fold_upper_elem(uint32_t w, uint32_t i, uint32_t l, uint32_t B, const XyzzPt<C> *buckets, const XyzzPt<C> *F, const uint32_t *bucket_offsets, XyzzPt<C> &out) {
    const uint32_t m = B >> l;
    if (l == 1) {
        if (bucket_offsets) {
            const uint32_t *off = bucket_offsets + (size_t)w * B + m;
            if (off[i] == off[i + 1]) return false;
        }
        out = buckets[(size_t)w * B + m + i];
    } else {
        out = F[(size_t)w * B + fold_level_offset(B, l - 1) + m + i];
    }
    return true;
}
// Term t of the window sum T0 + sum_b 2^b V_b: t < nb -> 2^t V_t, t == nb -> T0 = the single element of the last fold level.
template <class C> KGR_HD XyzzPt<C> fold_combine_term(uint32_t w, uint32_t t, uint32_t B, uint32_t nb, const XyzzPt<C> *F, const XyzzPt<C> *V) {
    if (t == nb) return F[(size_t)w * B + (B - 2)];
    XyzzPt<C> acc = V[(size_t)w * nb + t];
    for (uint32_t d = 0; d < t; d++) acc = xyzz_dbl(acc);
    return acc;
}

// Horner over the per-window sums (msm.rs:41,45-47 in one pass), then XYZZ -> (X : Y : Z).
template <class C> KGR_HD void body_final(const MsmShape &sh, const XyzzPt<C> *win_a, uint32_t *out24) {
    XyzzPt<C> r = win_a[sh.W - 1];
    for (uint32_t w = sh.W - 1; w-- > 0;) {
        for (uint32_t d = 0; d < sh.c; d++) r = xyzz_dbl(r);
        xyzz_add(r, win_a[w]);
    }
    typename C::Elem o[3];
    xyzz_to_projective(r, o);
    constexpr int NW = El<typename C::Elem>::WORDS;
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < NW; i++) out24[NW * k + i] = El<typename C::Elem>::word(o[k], i);
}

}  // namespace kgr
