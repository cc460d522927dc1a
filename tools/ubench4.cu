// Development microbenchmark (round 2): issue rate of the FP64 pipe on sm_100a (DFMA / DADD) and whether it overlaps the
// integer multiplier (IMAD.WIDE with carry) and the ALU pipe (IADD3) — in one warp's instruction stream and between
// warps of different kinds on the same SM sub-partition.  Rates per SASS instruction per clock per SMSP.
// Question behind it: can part of the Montgomery products of k_accumulate run as 52-bit double-precision products
// (two DFMA per 52x52 product, Emmart's technique) beside the IMAD.WIDE chains?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#define DECLS                                                                                 \
    uint32_t x[16];                                                                           \
    double d[8];                                                                              \
    for (int j = 0; j < 16; j++) x[j] = threadIdx.x * 7 + j * a;                              \
    for (int j = 0; j < 8; j++) d[j] = 1.0 + 1e-9 * (threadIdx.x + j);
#define FINISH                                                                                \
    uint32_t acc = 0;                                                                         \
    for (int j = 0; j < 16; j++) acc ^= x[j];                                                 \
    for (int j = 0; j < 8; j++) acc ^= (uint32_t)__double2loint(d[j]) ^ (uint32_t)__double2hiint(d[j]); \
    if (acc == 0x12345u) sink[0] = acc;
#define BODY_BEGIN(NAME)                                                                      \
    __global__ void __launch_bounds__(256) NAME(uint32_t *sink, uint32_t a, uint32_t b, double fa, double fb, int iters) { \
        DECLS                                                                                 \
        for (int it = 0; it < iters; it++) {                                                  \
            _Pragma("unroll") for (int rep = 0; rep < 8; rep++) {
#define BODY_END                                                                              \
            }                                                                                 \
        }                                                                                     \
        FINISH                                                                                \
    }

#define I_WIDE_CO(j) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(x[j]), "+r"(x[j + 8]) : "r"(a), "r"(b));
#define I_IADD(j) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(x[j + 8]));
#define I_IADD_B(j) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j + 8]) : "r"(x[(j + 1) & 7]));
#define I_ADD64(j) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(x[j]), "+r"(x[j + 8]) : "r"(x[(j + 1) & 7]), "r"(x[((j + 1) & 7) + 8]));
#define I_DFMA(j) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[j]) : "d"(fa), "d"(fb));
#define I_DFMA_RN(j) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[j]) : "d"(fa), "d"(fb));
#define I_DADD(j) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[j]) : "d"(fb));
#define I_DMUL(j) asm volatile("mul.rz.f64 %0, %0, %1;" : "+d"(d[j]) : "d"(fa));
#define I_LOP(j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[j + 8]), "r"(a));
// one 52x52 product the way the double-precision formulation does it: hi = fma_rz(a, b, c1); sub = c2 - hi; lo = fma_rz(a, b, sub);
// then two 64-bit integer accumulations of the raw bit patterns
#define I_EMMART(j)                                                                           \
    {                                                                                         \
        double hi, sub, lo;                                                                   \
        asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(hi) : "d"(d[j]), "d"(fa), "d"(fb));  \
        asm volatile("sub.rz.f64 %0, %1, %2;" : "=d"(sub) : "d"(fb), "d"(hi));                \
        asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(lo) : "d"(d[j]), "d"(fa), "d"(sub)); \
        uint32_t h0 = __double2loint(hi), h1 = __double2hiint(hi), l0 = __double2loint(lo), l1 = __double2hiint(lo); \
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(x[j]), "+r"(x[j + 8]) : "r"(h0), "r"(h1)); \
        asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(x[(j + 1) & 7]), "+r"(x[((j + 1) & 7) + 8]) : "r"(l0), "r"(l1)); \
        d[j] = lo; /* the next product of this slot depends on this one: ptxas must not merge the repetitions */ \
    }

BODY_BEGIN(k_wide_co) REP8(I_WIDE_CO) BODY_END
BODY_BEGIN(k_dfma) REP8(I_DFMA) BODY_END
BODY_BEGIN(k_dfma_rn) REP8(I_DFMA_RN) BODY_END
BODY_BEGIN(k_dadd) REP8(I_DADD) BODY_END
BODY_BEGIN(k_dmul) REP8(I_DMUL) BODY_END
BODY_BEGIN(k_dfma_wide) REP8(I_DFMA) REP8(I_WIDE_CO) BODY_END
BODY_BEGIN(k_dfma_iadd) REP8(I_DFMA) REP8(I_IADD) BODY_END
BODY_BEGIN(k_dfma_add64) REP8(I_DFMA) REP8(I_ADD64) BODY_END
BODY_BEGIN(k_dfma_lop) REP8(I_DFMA) REP8(I_LOP) BODY_END
BODY_BEGIN(k_dfma_wide_iadd) REP8(I_DFMA) REP8(I_WIDE_CO) REP8(I_IADD_B) BODY_END
BODY_BEGIN(k_emmart) REP8(I_EMMART) BODY_END
BODY_BEGIN(k_emmart_wide) REP8(I_EMMART) REP8(I_WIDE_CO) BODY_END

// warp-specialised: even warps run the integer multiplier, odd warps the FP64 pipe (n_fp of every 8 warps are FP64 warps)
__global__ void __launch_bounds__(256) k_spec(uint32_t *sink, uint32_t a, uint32_t b, double fa, double fb, int iters, int n_fp, int emmart) {
    DECLS
    int warp = threadIdx.x >> 5;
    if (warp < n_fp) {
        if (emmart) {
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int rep = 0; rep < 8; rep++) { REP8(I_EMMART) }
            }
        } else {
            for (int it = 0; it < iters; it++) {
#pragma unroll
                for (int rep = 0; rep < 8; rep++) { REP8(I_DFMA) }
            }
        }
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int rep = 0; rep < 8; rep++) { REP8(I_WIDE_CO) }
        }
    }
    FINISH
}

__global__ void k_clock(uint64_t *out) {
    uint64_t t0, t1, c0, c1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    c0 = clock64();
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < 2000000ULL);
    c1 = clock64();
    out[0] = t1 - t0; out[1] = c1 - c0;
}

typedef void (*kern_t)(uint32_t *, uint32_t, uint32_t, double, double, int);
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    uint32_t *sink; cudaMalloc(&sink, 4096);
    uint64_t *dclk, hclk[2]; cudaMalloc(&dclk, 16);
    const double fa = 1.0000000001, fb = 0.9999999999;
    struct { const char *name; kern_t k; int slots; } tests[] = {
        {"IMAD.WIDE+carry-out", k_wide_co, 8}, {"DFMA.RZ", k_dfma, 8}, {"DFMA.RN", k_dfma_rn, 8}, {"DADD", k_dadd, 8}, {"DMUL", k_dmul, 8},
        {"DFMA + WIDE.CO", k_dfma_wide, 16}, {"DFMA + IADD3", k_dfma_iadd, 16}, {"DFMA + add64 pair", k_dfma_add64, 24}, {"DFMA + LOP3", k_dfma_lop, 16},
        {"DFMA + WIDE.CO + IADD3", k_dfma_wide_iadd, 24}, {"emmart product (2 DFMA+DADD+4 IADD3)", k_emmart, 8}, {"emmart product + WIDE.CO", k_emmart_wide, 16}};
    auto clock_mhz = [&]() {
        k_clock<<<1, 1>>>(dclk);
        cudaDeviceSynchronize();
        cudaMemcpy(hclk, dclk, 16, cudaMemcpyDeviceToHost);
        return (double)hclk[1] / hclk[0] * 1e3;
    };
    for (int warps_per_smsp : {4, 16}) {
        int blocks = sms * warps_per_smsp / 2;  // 256 threads = 8 warps = 2 per SMSP
        printf("== %d warps per SMSP\n", warps_per_smsp);
        for (auto &t : tests) {
            const int iters = 1000;
            t.k<<<blocks, 256>>>(sink, 3, 5, fa, fb, 10);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0); t.k<<<blocks, 256>>>(sink, 3, 5, fa, fb, iters); cudaEventRecord(e1);
            double mhz = clock_mhz();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double slots = (double)warps_per_smsp * iters * 8.0 * t.slots;
            double cycles = ms * 1e-3 * mhz * 1e6;
            printf("%-40s %8.3f ms  %6.0f MHz  %.3f pattern-slots/clk/SMSP  (%.2f clk per slot)\n", t.name, ms, mhz, slots / cycles, cycles / slots);
        }
        // warp specialisation: report the clocks per instruction of each kind as if it ran alone in the same time
        for (int emmart = 0; emmart < 2; emmart++)
            for (int n_fp : {0, 2, 4, 6, 8}) {
                const int iters = 1000;
                k_spec<<<blocks, 256>>>(sink, 3, 5, fa, fb, 10, n_fp, emmart);
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0); k_spec<<<blocks, 256>>>(sink, 3, 5, fa, fb, iters, n_fp, emmart); cudaEventRecord(e1);
                double mhz = clock_mhz();
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                double cycles = ms * 1e-3 * mhz * 1e6;
                double fp_slots = (double)warps_per_smsp * n_fp / 8.0 * iters * 64.0, int_slots = (double)warps_per_smsp * (8 - n_fp) / 8.0 * iters * 64.0;
                printf("spec %s: %d of 8 warps FP64   %8.3f ms  FP slots/clk %.3f  WIDE slots/clk %.3f   (alone: FP %.3f, WIDE 0.248)\n", emmart ? "emmart" : "dfma  ", n_fp, ms,
                       fp_slots / cycles, int_slots / cycles, emmart ? 0.0 : 0.5);
            }
    }
    return 0;
}
