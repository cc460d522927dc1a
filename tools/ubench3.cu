// Development microbenchmark: per-instruction issue rates of the sm_100a integer pipes, and which pairs overlap.
// Rates are reported per SASS instruction per clock per SMSP, using the SM clock measured in-kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#define BODY_BEGIN(NAME)                                                                      \
    __global__ void __launch_bounds__(256) NAME(uint32_t *sink, uint32_t a, uint32_t b, int iters) { \
        uint32_t x[16];                                                                       \
        uint64_t y[8];                                                                        \
        for (int j = 0; j < 16; j++) x[j] = threadIdx.x * 7 + j * a;                          \
        for (int j = 0; j < 8; j++) y[j] = ((uint64_t)x[j] << 32) | x[j + 8];                 \
        for (int it = 0; it < iters; it++) {                                                  \
            _Pragma("unroll") for (int rep = 0; rep < 8; rep++) {
#define BODY_END                                                                              \
            }                                                                                 \
        }                                                                                     \
        uint32_t acc = 0;                                                                     \
        for (int j = 0; j < 16; j++) acc ^= x[j];                                             \
        for (int j = 0; j < 8; j++) acc ^= (uint32_t)y[j] ^ (uint32_t)(y[j] >> 32);           \
        if (acc == 0x12345u) sink[0] = acc;                                                   \
    }

// each "X(j)" is one instruction instance on independent registers
#define I_IMAD(j) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a), "r"(x[j + 8]));
#define I_IMADHI(j) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a), "r"(x[j + 8]));
#define I_WIDE(j) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(y[j]) : "r"(a), "r"(x[j]));
#define I_WIDE_CO(j) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(x[j]), "+r"(x[j + 8]) : "r"(a), "r"(b));
#define I_IADD(j) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(x[j + 8]));
#define I_IADD_B(j) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j + 8]) : "r"(x[(j + 1) & 7]));
#define I_IADDC(j) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(x[j]), "+r"(x[j + 8]) : "r"(a), "r"(b));
#define I_LOP(j) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(x[j + 8]), "r"(a));
#define I_SHF(j) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(x[j + 8]), "r"(b));
#define I_SHR64(j) y[j] = (y[j] >> 29) + x[j];
#define I_ADD64(j) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(x[j]), "+r"(x[j + 8]) : "r"(x[(j + 1) & 7]), "r"(x[((j + 1) & 7) + 8]));

BODY_BEGIN(k_imad) REP8(I_IMAD) BODY_END
BODY_BEGIN(k_imadhi) REP8(I_IMADHI) BODY_END
BODY_BEGIN(k_wide) REP8(I_WIDE) BODY_END
BODY_BEGIN(k_wide_co) REP8(I_WIDE_CO) BODY_END
BODY_BEGIN(k_iadd) REP8(I_IADD) BODY_END
BODY_BEGIN(k_iaddc) REP8(I_IADDC) BODY_END
BODY_BEGIN(k_lop) REP8(I_LOP) BODY_END
BODY_BEGIN(k_shf) REP8(I_SHF) BODY_END
BODY_BEGIN(k_shr64) REP8(I_SHR64) BODY_END
BODY_BEGIN(k_add64) REP8(I_ADD64) BODY_END
BODY_BEGIN(k_imad_iadd) REP8(I_IMAD) REP8(I_IADD_B) BODY_END
BODY_BEGIN(k_wide_iadd) REP8(I_WIDE) REP8(I_IADD_B) BODY_END
BODY_BEGIN(k_wide_lop) REP8(I_WIDE) REP8(I_LOP) BODY_END
BODY_BEGIN(k_wide_2iadd) REP8(I_WIDE) REP8(I_IADD_B) REP8(I_LOP) BODY_END
BODY_BEGIN(k_imad_wide) REP8(I_IMAD) REP8(I_WIDE) BODY_END

__global__ void k_clock(uint64_t *out) {
    uint64_t t0, t1, c0, c1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    c0 = clock64();
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < 2000000ULL);
    c1 = clock64();
    out[0] = t1 - t0; out[1] = c1 - c0;
}

typedef void (*kern_t)(uint32_t *, uint32_t, uint32_t, int);
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    uint32_t *sink; cudaMalloc(&sink, 4096);
    uint64_t *dclk, hclk[2]; cudaMalloc(&dclk, 16);
    struct { const char *name; kern_t k; int slots; } tests[] = {
        {"IMAD", k_imad, 8}, {"IMAD.HI", k_imadhi, 8}, {"IMAD.WIDE", k_wide, 8}, {"IMAD.WIDE+carry-out", k_wide_co, 8}, {"IADD3", k_iadd, 8},
        {"IADD3 cc pair", k_iaddc, 16}, {"LOP3", k_lop, 8}, {"SHF", k_shf, 8}, {"u64 (>>29)+u32", k_shr64, 8}, {"u64 add pair", k_add64, 16},
        {"IMAD + IADD3", k_imad_iadd, 16}, {"WIDE + IADD3", k_wide_iadd, 16}, {"WIDE + LOP3", k_wide_lop, 16}, {"WIDE + IADD3 + LOP3", k_wide_2iadd, 24},
        {"IMAD + WIDE", k_imad_wide, 16}};
    for (int warps_per_smsp : {4, 16}) {
        int blocks = sms * warps_per_smsp / 2;  // 256 threads = 8 warps = 2 per SMSP
        printf("== %d warps per SMSP\n", warps_per_smsp);
        for (auto &t : tests) {
            const int iters = 2000;
            t.k<<<blocks, 256>>>(sink, 3, 5, 10);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0); t.k<<<blocks, 256>>>(sink, 3, 5, iters); cudaEventRecord(e1);
            k_clock<<<1, 1>>>(dclk);
            cudaDeviceSynchronize();
            cudaMemcpy(hclk, dclk, 16, cudaMemcpyDeviceToHost);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double mhz = (double)hclk[1] / hclk[0] * 1e3;
            double warp_instrs_per_smsp = (double)warps_per_smsp * iters * 8.0 * t.slots;   // nominal pattern slots
            double cycles = ms * 1e-3 * mhz * 1e6;
            printf("%-22s %8.3f ms  %6.0f MHz  %.3f pattern-slots/clk/SMSP  (%.2f clk per slot)\n", t.name, ms, mhz, warp_instrs_per_smsp / cycles,
                   cycles / warp_instrs_per_smsp);
        }
    }
    return 0;
}
