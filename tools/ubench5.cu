// Development microbenchmark (round 2): what a random gather of 32-byte / 64-byte records from an array much larger than the L2 costs,
// under cudaLimitMaxL2FetchGranularity = 32 / 64 / 128 and with the ld.global.L2::64B / ::128B / ::256B prefetch-size qualifiers.
// Question behind it: the batched-affine level 0 (k_affine_den / k_affine_add) and k_accumulate gather 64-byte points by entry; ncu measured
// ~130 B of DRAM traffic per gathered point (the 128-byte line).  If the fetch granularity can be lowered to the record size, the level-0
// passes stop being DRAM-bound.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench5 tools/ubench5.cu ; run: tools/ubench5 [log2 records = 24]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));                  \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int MODE> __device__ __forceinline__ uint4 ld16(const uint4 *p) {
    uint4 v;
    if (MODE == 0) v = __ldg(p);
    else if (MODE == 1) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.nc.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// every thread gathers `per` records of REC bytes (REC = 32 or 64) at pseudo-random record indices; records are REC-aligned slots of 64 bytes
template <int REC, int MODE> __global__ void __launch_bounds__(256) k_gather(const uint4 *base, uint32_t mask, int per, uint32_t *sink) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    for (int i = 0; i < per; i += 4) {
        uint4 v[4][REC / 16];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            uint32_t r = mix(t * 131u + (uint32_t)(i + u) * 2654435761u) & mask;
#pragma unroll
            for (int q = 0; q < REC / 16; q++) v[u][q] = ld16<MODE>(base + (size_t)r * 4 + q);
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int q = 0; q < REC / 16; q++) acc ^= v[u][q].x ^ v[u][q].y ^ v[u][q].z ^ v[u][q].w;
    }
    if (acc == 0x1234567u) sink[0] = acc;
}

template <int REC, int MODE> static void run(const char *name, const uint4 *base, uint32_t mask, uint32_t *sink) {
    const int per = 64, blocks = 148 * 64, tpb = 256;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    k_gather<REC, MODE><<<blocks, tpb>>>(base, mask, per, sink);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    k_gather<REC, MODE><<<blocks, tpb>>>(base, mask, per, sink);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    double recs = (double)blocks * tpb * per;
    printf("  %-28s rec %2d B: %7.3f ms  %7.2f G records/s  %7.1f GB/s useful\n", name, REC, ms, recs / ms * 1e-6, recs * REC / ms * 1e-6);
}

int main(int argc, char **argv) {
    int logn = argc > 1 ? atoi(argv[1]) : 24;
    size_t n = (size_t)1 << logn;
    uint4 *base;
    uint32_t *sink;
    CK(cudaMalloc(&base, n * 64));
    CK(cudaMemset(base, 1, n * 64));
    CK(cudaMalloc(&sink, 4));
    uint32_t mask = (uint32_t)(n - 1);
    for (int g : {0, 32, 64, 128}) {
        if (g) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            printf("cudaLimitMaxL2FetchGranularity <- %d: %s, now %zu\n", g, cudaGetErrorString(e), got);
        } else {
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            printf("default cudaLimitMaxL2FetchGranularity = %zu; array %zu MiB of 64-byte slots\n", got, n * 64 >> 20);
        }
        run<64, 0>("ld.global.nc", base, mask, sink);
        run<64, 1>("ld.global.nc.L2::64B", base, mask, sink);
        run<64, 2>("ld.global.nc.L2::128B", base, mask, sink);
        run<64, 3>("ld.global.nc.L2::256B", base, mask, sink);
        run<64, 4>("ld.global.nc.L1::no_allocate", base, mask, sink);
        run<32, 0>("ld.global.nc", base, mask, sink);
        run<32, 1>("ld.global.nc.L2::64B", base, mask, sink);
        run<32, 2>("ld.global.nc.L2::128B", base, mask, sink);
    }
    return 0;
}
