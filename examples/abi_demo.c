/* examples/abi_demo.c — the C ABI of include/kgr_msm.h used from plain C (what a cgo / Rust `extern "C"` / JNI binding does), no Python.
 *
 *   gcc -O2 -I include -o abi_demo examples/abi_demo.c -L kogarashi_b200 -lkgr_msm -Wl,-rpath,$PWD/kogarashi_b200
 *
 * Without a CUDA device it must fail loudly (KGR_E_NO_DEVICE, exit code 3): there is no CPU fallback.  With one it checks, through the ABI only,
 *   msm([5 G, 7 G], [11, 13]) == (5 * 11 + 7 * 13) G = 146 G        on BN254 G1 and on BN254 G2,
 * once through kgr_msm_oneshot and once through kgr_bases_register + kgr_msm_batch; the multiples of the generator come from
 * kgr_fixed_base_mul. */
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
    int bad = check_curve(KGR_CURVE_BN254_G1, 4) | check_curve(KGR_CURVE_BN254_G2, 8);
    if (bad) printf("last error: %s\n", kgr_last_error());
    kgr_shutdown();
    return bad ? 1 : 0;
}
