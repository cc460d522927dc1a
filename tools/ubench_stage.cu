// Development microbenchmark: how fast can host threads copy pageable memory into a pinned buffer on this box (the staging step of
// upload_from_host for callers that pass ordinary memory — a Rust Vec, a numpy array)?  GB/s for 1 .. 16 threads, 96 MiB, 2-MiB chunks.
// Build: nvcc -O3 -o tools/ubench_stage tools/ubench_stage.cu
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
int main() {
    const size_t bytes = 96u << 20, chunk = 2u << 20;
    char *src = (char *)malloc(bytes), *dst;
    memset(src, 1, bytes);
    if (cudaMallocHost(&dst, bytes) != cudaSuccess) { printf("no device\n"); return 1; }
    memset(dst, 0, bytes);
    for (int nt : {1, 2, 4, 6, 8, 12, 16}) {
        double best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++)
                th.emplace_back([&, t]() {
                    for (size_t off = (size_t)t * chunk; off < bytes; off += (size_t)nt * chunk) memcpy(dst + off, src + off, chunk);
                });
            for (auto &x : th) x.join();
            best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        }
        printf("%2d threads: %.2f ms = %.1f GB/s (incl. thread start / join)\n", nt, best * 1e3, bytes / best / 1e9);
    }
    // one H2D copy of the pinned buffer for comparison
    void *d;
    cudaMalloc(&d, bytes);
    cudaMemcpy(d, dst, bytes, cudaMemcpyHostToDevice);
    auto t0 = std::chrono::steady_clock::now();
    cudaMemcpy(d, dst, bytes, cudaMemcpyHostToDevice);
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("pinned H2D: %.2f ms = %.1f GB/s\n", s * 1e3, bytes / s / 1e9);
    t0 = std::chrono::steady_clock::now();
    cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
    s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("pageable H2D (driver staging): %.2f ms = %.1f GB/s\n", s * 1e3, bytes / s / 1e9);
    return 0;
}
