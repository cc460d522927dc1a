// tests/host_emu.cpp — runs the MSM pipeline's per-thread bodies (kogarashi_b200/csrc/
// msm_kernels.cuh, compiled for the host: portable carry chains instead of PTX) thread by thread
// on the CPU and checks the result against the oracle (oracle/zkstd_oracle.hpp).  This validates
// the pipeline LOGIC (recoding, counting sort, chunked accumulation, fix-up, hierarchical
// reduction, Horner) without a GPU; the PTX bodies themselves are covered by the -m gpu tests.
//
// usage: host_emu <curve 0|1|2> <n> <c> <L> <K> <mode> <seed> [running_sum_stop]
//   mode 0 uniform scalars, 1 skewed (zeros/ones/r-1), 2 duplicate + opposite points + identity bases,
//        3 canonical-format scalars; add 10 for the window-collapsed (precomputed table) mode
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define KGR_HOST_EMULATE_CHAINS 1  // run the even/odd carry-chain algorithm, not the u128 host shortcut
#include "../kogarashi_b200/csrc/msm_kernels.cuh"
#include "../oracle/zkstd_oracle.hpp"

using namespace kgr;

template <class C, class OC> static int run(uint32_t n, uint32_t c, uint32_t L, uint32_t K, int mode, uint64_t seed, uint32_t rs_stop) {
    const bool collapsed = (mode >= 10);  // window-collapsed mode on a precomputed table
    if (collapsed) mode -= 10;
    typedef zko::Curve<OC> Cv;
    typedef zko::Field<typename OC::Scalar> Fs;
    typedef typename Cv::Affine OAffine;
    typedef typename Cv::Proj OProj;
    constexpr size_t EB = sizeof(typename C::Elem);  // bytes per coordinate: 32, or 64 for G2 (c0 || c1)
    static_assert(sizeof(typename Cv::El) == EB, "oracle and device coordinates have the same layout");
    std::mt19937_64 rng(seed);
    // points: small multiples of G built by repeated addition (cheap), shuffled
    std::vector<OAffine> pts(n);
    {
        OProj acc = Cv::to_extended(Cv::generator());
        OProj step = Cv::double_proj(acc);
        for (uint32_t i = 0; i < n; i++) {
            pts[i] = Cv::to_affine(acc);
            acc = Cv::add_proj(acc, step);
            if (rng() % 7 == 0) acc = Cv::double_proj(acc);
        }
    }
    std::vector<zko::Limbs> sc(n);
    for (uint32_t i = 0; i < n; i++) {
        uint64_t w[8];
        for (auto &x : w) x = rng();
        sc[i] = Fs::from_u512(w);
        if (mode == 1) {
            switch (rng() % 4) {
                case 0: sc[i] = Fs::zero(); break;
                case 1: sc[i] = Fs::one(); break;
                case 2: sc[i] = Fs::neg(Fs::one()); break;  // r - 1
                default: break;
            }
        }
    }
    if (mode == 2 && n >= 8) {
        for (uint32_t i = 0; i + 4 <= n; i += 4) {
            pts[i + 1] = pts[i];             // duplicate base
            pts[i + 2] = Cv::neg(pts[i]);    // opposite base
            sc[i + 1] = sc[i];               // same digits -> same buckets: P + P, then P - P paths
            sc[i + 2] = sc[i];
        }
        pts[3] = Cv::affine_identity();
        pts[n - 1] = Cv::affine_identity();
    }
    // expected (oracle, reference algorithm)
    OProj exp;
    {
        // inline copy of the reference structure through the naive sum: sum k_i P_i by scalar_point
        exp = Cv::proj_identity();
        for (uint32_t i = 0; i < n; i++) exp = Cv::add_proj(exp, Cv::scalar_point(Cv::to_extended(pts[i]), sc[i]));
    }
    OAffine exp_aff = Cv::to_affine(exp);

    // ---- device-format inputs
    std::vector<AffinePt<C>> bases(n);
    for (uint32_t i = 0; i < n; i++) {
        if (pts[i].inf) {
            memset(&bases[i], 0, sizeof bases[i]);
        } else {
            memcpy(&bases[i].x, &pts[i].x, EB);
            memcpy(&bases[i].y, &pts[i].y, EB);
        }
    }
    std::vector<uint32_t> scalars(8 * (size_t)n);
    int is_mont = (mode != 3);
    for (uint32_t i = 0; i < n; i++) {
        zko::Limbs v = is_mont ? sc[i] : Fs::montgomery_reduce(sc[i]);
        memcpy(&scalars[8 * (size_t)i], v.data(), 32);
    }
    MsmShape sh;
    sh.n = n; sh.c = c; sh.W = (255 + c - 1) / c; sh.B = 1u << (c - 1); sh.G = collapsed ? sh.B : sh.W * sh.B; sh.L = L; sh.K = K;
    sh.gstride = collapsed ? 0 : sh.B; sh.pstride = collapsed ? n + 5 : 0; sh.poff = collapsed ? 5 : 0;
    const uint32_t nwin = collapsed ? 1 : sh.W;
    std::vector<AffinePt<C>> table;
    if (collapsed) {  // table built over a longer vector (5 dummy leading points) to exercise pstride / poff
        std::vector<AffinePt<C>> ext(n + 5);
        for (uint32_t i = 0; i < 5; i++) ext[i] = bases[0];
        for (uint32_t i = 0; i < n; i++) ext[5 + i] = bases[i];
        table.resize((size_t)sh.W * (n + 5));
        for (uint32_t i = 0; i < n + 5; i++) body_precompute<C>(i, n + 5, sh.c, sh.W, n + 5, ext.data(), table.data());
    }
    const AffinePt<C> *acc_bases = collapsed ? table.data() : bases.data();
    // seed bit 5: the call arrives in two pieces (msm.cu: StreamPiece): each piece is sorted and accumulated on its own, its bucket sums are
    // merged into one shared, zero-initialised bucket set (body_bucket_merge), and one reduction over all buckets closes the call
    const uint32_t n_pieces = ((seed >> 5) & 1) && n >= 2 ? 2 : 1;
    const bool streamed = n_pieces > 1;
    const uint32_t chunks_max = ((uint64_t)n * sh.W + L - 1) / L + 1;
    std::vector<XyzzPt<C>> call_acc(sh.G), bucket_acc(sh.G), head(chunks_max), tail(chunks_max);
    memset(call_acc.data(), 0, call_acc.size() * sizeof(XyzzPt<C>));  // the device zero-fills the call's buckets: all-zero words are an XYZZ identity
    const uint32_t affine_levels = C::ID == 2 ? 0 : (seed >> 1) % 4;
    uint32_t M = 0;
    std::vector<uint32_t> offsets_last, off_cur_last;
    const MsmShape sh_call = sh;
    for (uint32_t piece = 0; piece < n_pieces; piece++) {
    const uint32_t first = piece ? n / 2 : 0, n_k = n_pieces == 1 ? n : (piece ? n - n / 2 : n / 2);
    MsmShape sh = sh_call;
    sh.n = n_k;
    if (collapsed) sh.poff = sh_call.poff + first;  // window table: same base pointer, shifted column
    const uint32_t *scalars_k = scalars.data() + 8 * (size_t)first;
    const AffinePt<C> *acc_bases_k = collapsed ? acc_bases : acc_bases + first;
    std::vector<uint32_t> counts(sh.G + 1, 0), offsets(sh.G + 2, 0);
    const bool window_major = (seed & 1) != 0;  // odd seeds exercise the window-major fill from stored digits
    std::vector<uint32_t> digits((size_t)n_k * sh.W + 1, 0x77777777u);
    for (uint32_t i = 0; i < n_k; i++) body_count<C>(i, sh, scalars_k, is_mont, counts.data(), window_major ? digits.data() : nullptr);
    uint32_t run_sum = 0;
    for (uint32_t g = 0; g <= sh.G; g++) { offsets[g] = run_sum; run_sum += counts[g]; }
    const uint32_t M_k = offsets[sh.G];
    M += M_k;
    std::vector<uint32_t> entries(M_k + 1, 0xdeadbeefu);
    if (window_major) {
        for (uint32_t w = 0; w < sh.W; w++)
            for (uint32_t i = 0; i < n_k; i++) body_fill_window<C>(i, w, sh, digits.data(), counts.data(), offsets.data(), entries.data());
    } else {
        for (uint32_t i = 0; i < n_k; i++) body_fill<C>(i, sh, scalars_k, is_mont, counts.data(), offsets.data(), entries.data());
    }
    for (uint32_t g = 0; g <= sh.G; g++) if (counts[g] != 0) { printf("FAIL counts not consumed at %u\n", g); return 1; }
    uint32_t chunks = ((uint64_t)n_k * sh.W + L - 1) / L + 1;
    // seed bit 6: the chunk length follows the entries actually present (eff_chunk_len): the grid is sized for the bound n * W, zero digits and
    // the batched-affine levels leave fewer entries, and accumulate / fix-up / hot-bucket sums must agree on the shorter chunks
    if ((seed >> 6) & 1) sh.chunks = chunks;
    // poison so that a missing write is noticed
    memset(bucket_acc.data(), 0xAB, bucket_acc.size() * sizeof(XyzzPt<C>));
    memset(head.data(), 0xCD, head.size() * sizeof(XyzzPt<C>));
    memset(tail.data(), 0xEF, tail.size() * sizeof(XyzzPt<C>));
    std::vector<uint32_t> tail_bucket(chunks, 0x12345678u);
    // seed bits 1-2: batched-affine tree levels in front of the XYZZ accumulation (affine_kernels.cuh).  The level structure — pairs by
    // position inside every bucket, the odd last node copied, off_{l+1} = scan of ceil(len / 2) — and the pair_case / pair_sum helpers are the
    // device code; the inverse of each denominator is computed directly here (the kernels share one inversion between many denominators by
    // Montgomery's trick, which yields the same field element).  G2 has no affine levels on the device.
    std::vector<uint32_t> off_cur(offsets.begin(), offsets.begin() + sh.G + 1);
    std::vector<AffinePt<C>> nodes;
    if constexpr (C::ID != 2) {
        for (uint32_t l = 0; l < affine_levels; l++) {
            std::vector<uint32_t> off_next(sh.G + 1, 0);
            for (uint32_t g = 0; g < sh.G; g++) off_next[g + 1] = off_next[g] + ((off_cur[g + 1] - off_cur[g] + 1) >> 1);
            std::vector<AffinePt<C>> next(off_next[sh.G]);
            auto load = [&](uint32_t pos) {
                if (l > 0) return nodes[pos];
                uint32_t ent = entries[pos];
                AffinePt<C> pt = acc_bases_k[ent & 0x7fffffffu];
                pt.y = fp_cneg(pt.y, (ent >> 31) != 0);
                return pt;
            };
            for (uint32_t g = 0; g < sh.G; g++)
                for (uint32_t q = off_next[g]; q < off_next[g + 1]; q++) {
                    uint32_t p0 = off_cur[g] + 2 * (q - off_next[g]);
                    AffinePt<C> a = load(p0), r = a;
                    if (p0 + 1 < off_cur[g + 1]) {
                        AffinePt<C> b = load(p0 + 1);
                        typename C::Elem den;
                        int code = pair_case(a, b, den);
                        r = pair_sum(code, a, b, fp_inv(den));
                    }
                    next[q] = r;
                }
            nodes.swap(next);
            off_cur.swap(off_next);
        }
    }
    const uint32_t *acc_off = affine_levels ? off_cur.data() : offsets.data();
    for (uint32_t t = 0; t < chunks; t++) {
        if (affine_levels)
            body_accumulate<C>(t, sh, nodes.data(), acc_off, (const uint32_t *)nullptr, bucket_acc.data(), head.data(), tail.data(), tail_bucket.data());
        else
            body_accumulate<C>(t, sh, acc_bases_k, offsets.data(), entries.data(), bucket_acc.data(), head.data(), tail.data(), tail_bucket.data());
    }
    std::vector<uint32_t> worklist(sh.G + 1);
    uint32_t wl_len = 0;
    for (uint32_t t = 0; t < chunks; t++) body_fixup<C>(t, sh, acc_off, bucket_acc.data(), head.data(), tail.data(), tail_bucket.data(), worklist.data(), &wl_len);
    for (uint32_t i = 0; i < wl_len; i++) {  // k_fixup_long: lanes cooperate, then a tree sum
        const uint32_t lanes = 5;
        XyzzPt<C> tot = xyzz_identity<C>();
        for (uint32_t lane = 0; lane < lanes; lane++) {
            XyzzPt<C> part = fixup_long_partial<C>(worklist[i], lane, lanes, sh, acc_off, head.data(), tail.data());
            xyzz_add(tot, part);
        }
        bucket_acc[worklist[i]] = tot;
    }
    if (streamed)
        for (uint32_t g = 0; g < sh.G; g++) body_bucket_merge<C>(g, sh.G, acc_off, bucket_acc.data(), call_acc.data());
    offsets_last = offsets;
    off_cur_last = off_cur;
    }  // pieces
    if (streamed) bucket_acc = call_acc;
    // the reduction: a streamed call reads every bucket (they are all defined), otherwise the empty ones are recognised from the offsets
    const uint32_t *acc_off = streamed ? nullptr : (affine_levels ? off_cur_last.data() : offsets_last.data());
    std::vector<XyzzPt<C>> win(nwin);
    const bool fold_reduce = ((seed >> 4) & 1) != 0 && sh.B >= 4;  // the default reduction of the GPU path (k_fold / k_fold_tail / k_vsum* / k_fold_combine)
    if (fold_reduce) {
        const uint32_t B = sh.B, nb = c - 1;
        std::vector<XyzzPt<C>> F((size_t)nwin * B), V((size_t)nwin * nb);
        memset(F.data(), 0x5A, F.size() * sizeof(XyzzPt<C>));
        for (uint32_t l = 1; l <= nb; l++)
            for (uint32_t w = 0; w < nwin; w++)
                for (uint32_t i = 0; i < (B >> l); i++) body_fold<C>(w, i, l, B, bucket_acc.data(), F.data(), acc_off);
        for (uint32_t l = 1; l <= nb; l++)
            for (uint32_t w = 0; w < nwin; w++) {
                XyzzPt<C> acc = xyzz_identity<C>(), x;
                for (uint32_t i = 0; i < (B >> l); i++)
                    if (fold_upper_elem<C>(w, i, l, B, bucket_acc.data(), F.data(), acc_off, x)) xyzz_add(acc, x);
                V[(size_t)w * nb + (nb - l)] = acc;
            }
        for (uint32_t w = 0; w < nwin; w++) {
            win[w] = xyzz_identity<C>();
            for (uint32_t t = 0; t <= nb; t++) {
                XyzzPt<C> term = fold_combine_term<C>(w, t, B, nb, F.data(), V.data());
                xyzz_add(win[w], term);
            }
        }
    } else {
    uint32_t cnt = sh.B, m_log2 = 0, klog = 0;
    while ((1u << klog) < K) klog++;
    uint32_t cnt1 = (sh.B + K - 1) / K;
    std::vector<XyzzPt<C>> ls[2], la[2];
    for (int i = 0; i < 2; i++) { ls[i].resize((size_t)nwin * cnt1); la[i].resize((size_t)nwin * cnt1); }
    const XyzzPt<C> *in_s = bucket_acc.data(), *in_a = nullptr;
    int pp = 0;
    for (;;) {
        uint32_t cnt_out = (cnt + K - 1) / K;
        for (uint32_t t = 0; t < nwin * cnt_out; t++) body_reduce<C>(t, nwin, cnt, K, m_log2, in_s, in_a, ls[pp].data(), la[pp].data(), in_a ? nullptr : acc_off);
        in_s = ls[pp].data(); in_a = la[pp].data();
        cnt = cnt_out; m_log2 += klog; pp ^= 1;
        if (cnt <= rs_stop) break;
    }
    if (cnt > 1) {  // k_weight + k_tree_sum
        std::vector<XyzzPt<C>> v((size_t)nwin * cnt);
        for (uint32_t t = 0; t < nwin * cnt; t++) body_weight<C>(t, nwin, cnt, m_log2, in_s, in_a, v.data());
        for (uint32_t w = 0; w < nwin; w++) {
            win[w] = xyzz_identity<C>();
            for (uint32_t i = 0; i < cnt; i++) xyzz_add(win[w], v[(size_t)w * cnt + i]);
        }
    } else {
        for (uint32_t w = 0; w < nwin; w++) win[w] = in_a[w];
    }
    }
    uint32_t out24[48];
    MsmShape shf = sh;
    shf.W = nwin;  // collapsed: a single window sum, no doublings left
    body_final<C>(shf, win.data(), out24);
    OProj got;
    memcpy(&got.x, out24, EB); memcpy(&got.y, out24 + EB / 4, EB); memcpy(&got.z, out24 + 2 * (EB / 4), EB);
    OAffine got_aff = Cv::to_affine(got);
    bool ok = Cv::eq(got_aff, exp_aff) && (got_aff.inf || (got_aff.x == exp_aff.x && got_aff.y == exp_aff.y));
    printf("%s curve=%d n=%u c=%u W=%u L=%u K=%u mode=%d M=%u inf=%d affine_levels=%u fold=%d\n", ok ? "OK" : "FAIL", C::ID, n, c, sh.W, L, K, mode, M, (int)got_aff.inf, affine_levels, (int)fold_reduce);
    return ok ? 0 : 1;
}

int main(int argc, char **argv) {
    if (argc < 8) { fprintf(stderr, "usage: host_emu curve n c L K mode seed\n"); return 2; }
    int curve = atoi(argv[1]);
    uint32_t n = (uint32_t)atol(argv[2]), c = (uint32_t)atol(argv[3]), L = (uint32_t)atol(argv[4]), K = (uint32_t)atol(argv[5]);
    int mode = atoi(argv[6]);
    uint64_t seed = strtoull(argv[7], nullptr, 10);
    uint32_t rs_stop = argc > 8 ? (uint32_t)atol(argv[8]) : 1;
    if (curve == 0) return run<Bn254G1, zko::Bn254G1>(n, c, L, K, mode, seed, rs_stop);
    if (curve == 2) return run<Bn254G2, zko::Bn254G2>(n, c, L, K, mode, seed, rs_stop);
    return run<GrumpkinC, zko::Grumpkin>(n, c, L, K, mode, seed, rs_stop);
}
