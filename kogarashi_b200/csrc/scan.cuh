// scan.cuh — hand-written exclusive prefix sum over uint32 (no CUB/Thrust).
//
// Three launches: tile scan (warp shuffles + one shared-memory hop, 4096 elements per CTA, 128-bit
// loads/stores), a single-CTA scan of the tile totals, and a uniform add.  The arrays scanned here
// are the bucket histograms (W * 2^(c-1) + 1 counters, <= 28 MB), so the scan is a few tens of
// microseconds and L2 resident; it is not worth a decoupled look-back.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace kgr {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *smem /* 33 words */) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < (blockDim.x >> 5)) ? smem[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
            if (lane >= (unsigned)d) winc += t;
        }
        smem[lane] = winc - w;  // exclusive prefix of warp totals
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    uint32_t r = smem[warp] + inc - v;
    *total = smem[32];
    __syncthreads();
    return r;
}

// in/out may alias.  len need not be a multiple of anything; in/out must be 16-byte aligned.
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t *in, uint32_t *out, uint32_t len, uint32_t *tile_sums) {
    __shared__ uint32_t smem[33];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    if (base + SCAN_ITEMS <= len) {
        uint4 q = *reinterpret_cast<const uint4 *>(in + base);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) v[i] = (base + i < len) ? in[base + i] : 0u;
    }
    uint32_t sum = v[0] + v[1] + v[2] + v[3];
    uint32_t total;
    uint32_t ex = block_exclusive_scan(sum, &total, smem);
    uint32_t o[SCAN_ITEMS];
    o[0] = ex; o[1] = ex + v[0]; o[2] = o[1] + v[1]; o[3] = o[2] + v[2];
    if (base + SCAN_ITEMS <= len) {
        *reinterpret_cast<uint4 *>(out + base) = make_uint4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++)
            if (base + i < len) out[base + i] = o[i];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of tile_sums[0..n) in place
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(uint32_t *tile_sums, uint32_t n) {
    __shared__ uint32_t smem[33];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += SCAN_THREADS) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = (i < n) ? tile_sums[i] : 0u;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, &total, smem);
        if (i < n) tile_sums[i] = ex + carry;
        carry += total;
    }
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t *out, uint32_t len, const uint32_t *tile_sums) {
    uint32_t add = tile_sums[blockIdx.x];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    if (base + SCAN_ITEMS <= len) {
        uint4 q = *reinterpret_cast<uint4 *>(out + base);
        q.x += add; q.y += add; q.z += add; q.w += add;
        *reinterpret_cast<uint4 *>(out + base) = q;
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++)
            if (base + i < len) out[base + i] += add;
    }
}

inline uint32_t scan_num_tiles(uint32_t len) { return (len + SCAN_TILE - 1) / SCAN_TILE; }

// out[i] = sum_{j<i} in[j];  tile_sums must hold scan_num_tiles(len) words.
inline void exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint32_t len, uint32_t *tile_sums, cudaStream_t st) {
    uint32_t tiles = scan_num_tiles(len);
    k_scan_tiles<<<tiles, SCAN_THREADS, 0, st>>>(in, out, len, tile_sums);
    if (tiles > 1) {
        k_scan_sums<<<1, SCAN_THREADS, 0, st>>>(tile_sums, tiles);
        k_scan_add<<<tiles, SCAN_THREADS, 0, st>>>(out, len, tile_sums);
    }
}

}  // namespace kgr
