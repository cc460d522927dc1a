/* examples/abi_demo.c — the C ABI of include/kgr_msm.h used from plain C (what a cgo / Rust `extern "C"` / JNI binding does), no Python.
 *
 *   gcc -O2 -I include -o abi_demo examples/abi_demo.c -L kogarashi_b200 -lkgr_msm -Wl,-rpath,$PWD/kogarashi_b200
 *
 * Without a CUDA device it must fail loudly (KGR_E_NO_DEVICE, exit code 3): there is no CPU fallback.  With one it checks, through the ABI only,
 *   msm([5 G, 7 G], [11, 13]) == (5 * 11 + 7 * 13) G = 146 G        on BN254 G1 and on BN254 G2,
 * once through kgr_msm_oneshot and once through kgr_bases_register + kgr_msm_batch; the multiples of the generator come from
 * kgr_fixed_base_mul.  check_marshal() then walks through the steps of the Rust crate (rust/kogarashi-msm-b200/src/lib.rs) one by one:
 * points and scalars packed into a reusable pinned arena (PinnedArena / kgr_host_alloc), is_identity flags, Montgomery scalars as they sit
 * in `Fr.0`, a registered vector used with a base offset (`&params.a[l..]`), and a commitment from a vector that lives on the GPU
 * (DeviceVec / kgr_vec_upload + kgr_msm_vec) — every result is compared with the host-buffer call. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "kgr_msm.h"

static int same_affine(const uint64_t *a, const uint64_t *b, int words) { return memcmp(a, b, (size_t)words * 8) == 0; }

static int check_curve(int curve, int coord_limbs) {
    /* kgr_fixed_base_mul takes Montgomery scalars: convert 5, 7 and 146 with the library's field hook (Fr, op 6 = to_mont) */
    uint64_t mont[3][4];
    uint64_t k_c[3][4] = {{5, 0, 0, 0}, {7, 0, 0, 0}, {5 * 11 + 7 * 13, 0, 0, 0}};
    for (int i = 0; i < 3; i++)
        if (kgr_test_field_op(1 /* Fr */, 6 /* to_mont */, k_c[i], NULL, 1, mont[i])) return 1;
    uint64_t pts[3][16];
    if (kgr_fixed_base_mul(curve, &mont[0][0], 3, &pts[0][0])) return 1; /* 5G, 7G, 146G, packed x||y */
    uint64_t packed[3 * 16];
    memcpy(packed, pts, sizeof packed);
    const int pw = 2 * coord_limbs; /* words per affine point */
    uint64_t s_c[2][4] = {{11, 0, 0, 0}, {13, 0, 0, 0}};
    uint64_t proj[24], aff[17];
    if (kgr_msm_oneshot(curve, packed, NULL, 2, &s_c[0][0], KGR_SCALARS_CANONICAL, 2, proj)) return 1;
    if (kgr_to_affine(curve, proj, aff)) return 1;
    int ok = aff[pw] == 0 && same_affine(aff, packed + 2 * pw, pw);
    /* registered bases + batch of two jobs give the same point */
    kgr_bases_t *b = NULL;
    if (kgr_bases_register(curve, packed, NULL, 2, &b)) return 1;
    uint64_t p1[24], p2[24];
    kgr_msm_job_t jobs[2] = {{b, 0, &s_c[0][0], KGR_SCALARS_CANONICAL, 2, p1}, {b, 1, &s_c[1][0], KGR_SCALARS_CANONICAL, 1, p2}};
    if (kgr_msm_batch(jobs, 2)) return 1;
    if (kgr_to_affine(curve, p1, aff)) return 1;
    ok = ok && aff[pw] == 0 && same_affine(aff, packed + 2 * pw, pw);
    kgr_bases_free(b);
    printf("curve %d: msm([5G, 7G], [11, 13]) == 146 G : %s\n", curve, ok ? "yes" : "NO");
    return ok ? 0 : 1;
}

/* The Rust shim's marshalling, step by step, on BN254 G1 with n = 4 points (the third one an identity). */
static int check_marshal(void) {
    enum { N = 4 };
    uint64_t k_c[N][4] = {{3, 0, 0, 0}, {9, 0, 0, 0}, {1, 0, 0, 0}, {12345, 0, 0, 0}}, k_m[N][4];
    uint64_t s_c[N][4] = {{2, 0, 0, 0}, {5, 0, 0, 0}, {77, 0, 0, 0}, {1000003, 0, 0, 0}}, s_m[N][4];
    for (int i = 0; i < N; i++)
        if (kgr_test_field_op(1, 6, k_c[i], NULL, 1, k_m[i]) || kgr_test_field_op(1, 6, s_c[i], NULL, 1, s_m[i])) return 1;
    /* PinnedArena::words(): one page-locked buffer for x || y of every point followed by the flag bytes, another for the scalars */
    void *arena_p = NULL, *arena_s = NULL;
    if (kgr_host_alloc(N * 8 * sizeof(uint64_t) + N, &arena_p) || kgr_host_alloc(N * 4 * sizeof(uint64_t), &arena_s)) return 1;
    uint64_t *xy = (uint64_t *)arena_p;
    uint8_t *inf = (uint8_t *)(xy + N * 8);
    uint64_t *sc = (uint64_t *)arena_s;
    if (kgr_fixed_base_mul(KGR_CURVE_BN254_G1, &k_m[0][0], N, xy)) return 1; /* GpuCurve::write_coords: x.inner() || y.inner() */
    memset(inf, 0, N);
    inf[2] = 1;                                    /* b.is_identity() as u8: that base must not contribute */
    memcpy(sc, s_m, sizeof s_m);                   /* GpuCurve::scalar_limbs: Fr.0, Montgomery form */
    uint64_t proj[12], a1[9], a2[9], a3[9], a4[9];
    if (kgr_msm_oneshot(KGR_CURVE_BN254_G1, xy, inf, N, sc, KGR_SCALARS_MONTGOMERY, N, proj) || kgr_to_affine(KGR_CURVE_BN254_G1, proj, a1)) return 1;
    /* expected: (3*2 + 9*5 + 12345*1000003) G through the fixed-base path */
    uint64_t e_c[4] = {3 * 2 + 9 * 5 + 12345ull * 1000003ull, 0, 0, 0}, e_m[4], e_xy[8];
    if (kgr_test_field_op(1, 6, e_c, NULL, 1, e_m) || kgr_fixed_base_mul(KGR_CURVE_BN254_G1, e_m, 1, e_xy)) return 1;
    int ok = a1[8] == 0 && same_affine(a1, e_xy, 8);
    /* RegisteredBases::new + msm(offset, ..): the tail [1..] of the vector with the tail of the scalars */
    kgr_bases_t *b = NULL;
    if (kgr_bases_register(KGR_CURVE_BN254_G1, xy, inf, N, &b)) return 1;
    if (kgr_msm(b, 0, sc, KGR_SCALARS_MONTGOMERY, N, proj) || kgr_to_affine(KGR_CURVE_BN254_G1, proj, a2)) return 1;
    ok = ok && same_affine(a1, a2, 9);
    uint64_t head[12], tail[12], sum[12];
    if (kgr_msm(b, 0, sc, KGR_SCALARS_MONTGOMERY, 1, head) || kgr_msm(b, 1, sc + 4, KGR_SCALARS_MONTGOMERY, N - 1, tail)) return 1;
    if (kgr_proj_add(KGR_CURVE_BN254_G1, head, tail, sum) || kgr_to_affine(KGR_CURVE_BN254_G1, sum, a3)) return 1;
    ok = ok && same_affine(a1, a3, 9);
    /* DeviceVec::upload + RegisteredBases::commit_resident: the scalars never leave the GPU between the two calls */
    kgr_vec_t *v = NULL;
    if (kgr_vec_upload(1 /* Fr */, sc, N, &v)) return 1;
    if (kgr_pedersen_commit_vec(b, v, 0, N, a4)) return 1;
    ok = ok && same_affine(a1, a4, 9) && kgr_vec_len(v) == N;
    kgr_vec_free(v);
    kgr_bases_free(b);
    kgr_host_free(arena_p);
    kgr_host_free(arena_s);
    printf("marshal steps of the Rust shim (pinned arena, flags, offsets, resident scalars) agree : %s\n", ok ? "yes" : "NO");
    return ok ? 0 : 1;
}

int main(void) {
    int rc = kgr_init(NULL, 0);
    if (rc == KGR_E_NO_DEVICE) {
        printf("kgr_init: %s (no CPU fallback)\n", kgr_last_error());
        return 3;
    }
    if (rc) {
        printf("kgr_init failed: %d %s\n", rc, kgr_last_error());
        return 2;
    }
    int bad = check_curve(KGR_CURVE_BN254_G1, 4) | check_curve(KGR_CURVE_BN254_G2, 8) | check_marshal();
    if (bad) printf("last error: %s\n", kgr_last_error());
    kgr_shutdown();
    return bad ? 1 : 0;
}
