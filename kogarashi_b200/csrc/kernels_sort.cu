// kernels_sort.cu — bucket sort of the (scalar, window) digits by two block-local radix partitions (round 2).
//
// Replaces the scatter of groth16/src/msm.rs:25-30 (every scalar walks to its bucket) for large inputs.  The round-1 counting sort
// (k_count / k_fill in msm_kernels.cuh, still used for small inputs) pays one L2 atomic and one scattered 4-byte store per
// (scalar, window); here every global access is a run of consecutive addresses and every atomic is a shared-memory one:
//
//   k_sort_digits     one thread per scalar: canonical form (one Montgomery product), signed c-bit recoding, digit words stored
//                     window-major (coalesced); histogram of the COARSE bins (top bits of the bucket index) of every window in
//                     shared memory, flushed once per CTA
//   k_sort_scan       one CTA: exclusive scan of the nseg * nbins coarse counters -> bin offsets + a cursor copy; counters zeroed again
//   k_sort_partition  CTA = (tile of T scalars, window): ranks its digits by coarse bin in shared memory, claims one range per bin
//                     with ONE global atomic, writes (payload, fine byte) runs of ~T / nbins consecutive elements
//   k_sort_buckets    CTA = one coarse bin: histogram of the fine bits (<= 256 buckets) -> offsets[] of its buckets; elements placed
//                     in shared memory by bucket, written out linearly -> entries[]
//
// Afterwards entries[] / offsets[] have exactly the meaning the round-1 sort gives them (offsets[g] exclusive, offsets[G] = total),
// so accumulate / fixup / reduce are unchanged.  The order of the entries inside one bucket depends on the order in which tiles claim
// their ranges; a bucket's sum does not.
//
// A segment is an independent bucket set: one per window, or a single one when the bases are a precomputed window table
// (MsmShape::gstride == 0: all windows share the buckets).
#include <cuda_runtime.h>

#include <algorithm>

#include "launch.cuh"
#include "scan.cuh"

namespace kgr {

constexpr int SORT_TPB = 1024;

template <class P>
__global__ void __launch_bounds__(SORT_TPB) k_sort_digits(SortPlan pl, const uint32_t *scalars, int is_mont, uint32_t *digits, uint32_t *coarse_counts) {
    extern __shared__ uint32_t hist[];  // nseg * nbins
    const uint32_t nh = pl.nseg * pl.nbins;
    for (uint32_t k = threadIdx.x; k < nh; k += SORT_TPB) hist[k] = 0;
    __syncthreads();
    const uint32_t B = 1u << (pl.c - 1);
    for (uint32_t i = blockIdx.x * SORT_TPB + threadIdx.x; i < pl.n; i += gridDim.x * SORT_TPB) {
        Fp<P> s;
        const uint4 *p = reinterpret_cast<const uint4 *>(scalars) + 2 * (size_t)i;
        uint4 lo = p[0], hi = p[1];
        s.v[0] = lo.x; s.v[1] = lo.y; s.v[2] = lo.z; s.v[3] = lo.w;
        s.v[4] = hi.x; s.v[5] = hi.y; s.v[6] = hi.z; s.v[7] = hi.w;
        // Montgomery input: the product with 1 reduces any 256-bit value; canonical input: at most five subtractions of the modulus
        if (is_mont) s = fp_from_mont(s);
        else reduce_range(s);
        // Digits are cut from a 64-bit bit buffer refilled limb by limb, so that every limb index is a compile-time constant (a window
        // loop indexing s.v[bit >> 5] costs ~60 instructions per window in register selects: 1400 per scalar measured, issue-bound).
        uint64_t buf = 0;
        uint32_t have = 0, w = 0, carry = 0;
        const uint32_t mask = (1u << pl.c) - 1u;
        auto emit = [&](uint32_t raw) {
            uint32_t d = raw + carry, sign = 0;
            carry = 0;
            if (d > B) {
                d = (1u << pl.c) - d;
                sign = 1;
                carry = 1;
            }
            digits[(size_t)w * pl.n + i] = d ? ((d - 1) | (sign << 31)) : NO_DIGIT;
            if (d) atomicAdd(&hist[(pl.nseg > 1 ? w : 0) * pl.nbins + ((d - 1) >> pl.F)], 1u);
            w++;
        };
#pragma unroll
        for (int l = 0; l < 8; l++) {
            buf |= (uint64_t)s.v[l] << have;
            have += 32;
            while (have >= pl.c && w < pl.W) {
                emit((uint32_t)buf & mask);
                buf >>= pl.c;
                have -= pl.c;
            }
        }
        while (w < pl.W) {  // the top window holds the bits that are left (fewer than c)
            emit((uint32_t)buf & mask);
            buf >>= pl.c;
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < nh; k += SORT_TPB) {
        uint32_t v = hist[k];
        if (v) atomicAdd(&coarse_counts[k], v);
    }
}

// one CTA: coarse_off[k] = exclusive sum (k = 0 .. nh, so coarse_off[nh] = total), cursor = copy, counts zeroed for the next MSM,
// offsets_total[0] = total (offsets[G] of the bucket array)
__global__ void __launch_bounds__(SORT_TPB) k_sort_scan(uint32_t nh, uint32_t *coarse_counts, uint32_t *coarse_off, uint32_t *cursor, uint32_t *offsets_total) {
    __shared__ uint32_t smem[33];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nh; base += SORT_TPB) {
        uint32_t k = base + threadIdx.x;
        uint32_t v = (k < nh) ? coarse_counts[k] : 0u;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, &total, smem);
        if (k < nh) {
            coarse_off[k] = ex + carry;
            cursor[k] = ex + carry;
            coarse_counts[k] = 0;
        }
        carry += total;
    }
    if (threadIdx.x == 0) {
        coarse_off[nh] = carry;
        offsets_total[0] = carry;
    }
}

// dynamic shared memory: cnt[nbins] | lstart[nbins + 1] | gbase[nbins] | stage_pay[T] | stage_fine[T] (bytes)
// The tile is read twice (the second time from L1 / L2) and ranked with two shared-memory atomics per element.  Keeping the words and the
// ranks of the first pass in registers instead was measured slower (51 registers: one CTA per SM instead of two, 1.17 vs 1.06 ms at 2^24):
// the kernel is a sequence of barrier-separated phases and needs the second CTA to fill them.
__global__ void __launch_bounds__(SORT_TPB) k_sort_partition(SortPlan pl, const uint32_t *digits, uint32_t *cursor, uint32_t *part_pay, uint8_t *part_fine) {
    extern __shared__ uint32_t sm[];
    uint32_t *cnt = sm, *lstart = cnt + pl.nbins, *gbase = lstart + pl.nbins + 1, *stage_pay = gbase + pl.nbins;
    uint8_t *stage_fine = reinterpret_cast<uint8_t *>(stage_pay + pl.T);
    __shared__ uint32_t scan_sm[33];
    const uint32_t w = blockIdx.y, i0 = blockIdx.x * pl.T;
    const uint32_t count = min(pl.T, pl.n - i0);
    const uint32_t seg = pl.nseg > 1 ? w : 0;
    const uint32_t *src = digits + (size_t)w * pl.n + i0;
    const uint32_t fmask = (1u << pl.F) - 1u, pay0 = w * pl.pstride + pl.poff + i0;
    for (uint32_t k = threadIdx.x; k < pl.nbins; k += SORT_TPB) cnt[k] = 0;
    __syncthreads();
    {
        for (uint32_t k0 = 0; k0 < count; k0 += 4 * SORT_TPB) {
            uint32_t wd[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                uint32_t k = k0 + u * SORT_TPB + threadIdx.x;
                wd[u] = k < count ? src[k] : NO_DIGIT;
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (wd[u] != NO_DIGIT) atomicAdd(&cnt[(wd[u] & 0x7fffffffu) >> pl.F], 1u);
        }
    }
    __syncthreads();
    // exclusive scan of cnt[] -> lstart[]; every thread owns `per` consecutive bins and claims their global ranges
    {
        const uint32_t per = (pl.nbins + SORT_TPB - 1) / SORT_TPB;
        uint32_t b0 = threadIdx.x * per, sum = 0;
        for (uint32_t j = 0; j < per; j++)
            if (b0 + j < pl.nbins) sum += cnt[b0 + j];
        uint32_t total;
        uint32_t ex = block_exclusive_scan(sum, &total, scan_sm);
        for (uint32_t j = 0; j < per; j++) {
            uint32_t b = b0 + j;
            if (b < pl.nbins) {
                uint32_t c = cnt[b];
                lstart[b] = ex;
                gbase[b] = c ? atomicAdd(&cursor[seg * pl.nbins + b], c) : 0u;
                cnt[b] = ex;  // becomes the running staging position of the second pass (start + rank: no second lookup of lstart[])
                ex += c;
            }
        }
        if (threadIdx.x == 0) lstart[pl.nbins] = total;
    }
    __syncthreads();
    {
        for (uint32_t k0 = 0; k0 < count; k0 += 4 * SORT_TPB) {
            uint32_t wd[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                uint32_t k = k0 + u * SORT_TPB + threadIdx.x;
                wd[u] = k < count ? src[k] : NO_DIGIT;  // second read: L1 / L2 hit
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (wd[u] != NO_DIGIT) {
                    uint32_t b = wd[u] & 0x7fffffffu, bin = b >> pl.F;
                    uint32_t pos = atomicAdd(&cnt[bin], 1u);
                    stage_pay[pos] = (pay0 + k0 + u * SORT_TPB + threadIdx.x) | (wd[u] & 0x80000000u);
                    stage_fine[pos] = (uint8_t)(b & fmask);
                }
            }
        }
    }
    __syncthreads();
    // write-out: groups of 16 lanes copy one bin's run each (a run is ~T / nbins >= 8 consecutive elements on both sides)
    for (uint32_t bin = threadIdx.x >> 4; bin < pl.nbins; bin += SORT_TPB >> 4) {
        const uint32_t lo = lstart[bin], hi = lstart[bin + 1], g0 = gbase[bin];
        for (uint32_t pos = lo + (threadIdx.x & 15); pos < hi; pos += 16) {
            part_pay[g0 + (pos - lo)] = stage_pay[pos];
            part_fine[g0 + (pos - lo)] = stage_fine[pos];
        }
    }
}

// dynamic shared memory: sorted[cap]
__global__ void __launch_bounds__(SORT_TPB) k_sort_buckets(SortPlan pl, const uint32_t *coarse_off, const uint32_t *part_pay, const uint8_t *part_fine,
                                                          uint32_t *entries, uint32_t *offsets) {
    extern __shared__ uint32_t sorted[];
    __shared__ uint32_t cnt[256], scan_sm[33];
    const uint32_t nf = 1u << pl.F;
    for (uint32_t sb = blockIdx.x; sb < pl.nseg * pl.nbins; sb += gridDim.x) {
        const uint32_t base = coarse_off[sb], size = coarse_off[sb + 1] - base;
        const uint32_t seg = sb / pl.nbins, bin = sb % pl.nbins;
        if (threadIdx.x < 256) cnt[threadIdx.x] = 0;
        __syncthreads();
        {
            for (uint32_t k0 = 0; k0 < size; k0 += 8 * SORT_TPB) {
                uint32_t f[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    uint32_t k = k0 + u * SORT_TPB + threadIdx.x;
                    f[u] = k < size ? part_fine[base + k] : 0xffffffffu;
                }
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (f[u] != 0xffffffffu) atomicAdd(&cnt[f[u]], 1u);
            }
        }
        __syncthreads();
        {
            uint32_t v = threadIdx.x < nf ? cnt[threadIdx.x] : 0u, total;
            uint32_t ex = block_exclusive_scan(v, &total, scan_sm);
            if (threadIdx.x < nf) {
                cnt[threadIdx.x] = ex;  // running position of the placement pass (bucket start + rank)
                offsets[(size_t)seg * (pl.nbins << pl.F) + ((size_t)bin << pl.F) + threadIdx.x] = base + ex;
            }
        }
        __syncthreads();
        if (size <= pl.cap) {
            for (uint32_t k0 = 0; k0 < size; k0 += 8 * SORT_TPB) {
                uint32_t f[8], py[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    uint32_t k = k0 + u * SORT_TPB + threadIdx.x;
                    f[u] = k < size ? part_fine[base + k] : 0xffffffffu;
                    py[u] = k < size ? part_pay[base + k] : 0u;
                }
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (f[u] != 0xffffffffu) sorted[atomicAdd(&cnt[f[u]], 1u)] = py[u];
            }
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < size; k += SORT_TPB) entries[base + k] = sorted[k];
        } else {
            // a bin that outgrew the shared-memory buffer (skewed scalars, or the thin top window): same placement, written straight to global
            for (uint32_t k = threadIdx.x; k < size; k += SORT_TPB) {
                uint32_t f = part_fine[base + k];
                entries[base + atomicAdd(&cnt[f], 1u)] = part_pay[base + k];
            }
        }
        __syncthreads();
    }
}

// ---- plan + launchers ---------------------------------------------------------------------------------------------------------------
static int g_smem_optin = -1;
static int smem_optin() {
    if (g_smem_optin < 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        g_smem_optin = v;
    }
    return g_smem_optin;
}

bool LaunchSort::plan(const MsmShape &sh, SortPlan &pl) {
    pl.n = sh.n; pl.c = sh.c; pl.W = sh.W;
    pl.nseg = sh.gstride ? sh.W : 1;
    pl.pstride = sh.pstride; pl.poff = sh.poff;
    if (sh.c < 2) return false;
    const uint32_t bits = sh.c - 1;                       // bucket index bits
    const uint64_t per_seg = (uint64_t)sh.n * sh.W / pl.nseg;  // expected elements per segment (at most)
    // coarse bins: enough that a bin (per_seg / nbins elements for uniform digits) fits the shared-memory buffer of k_sort_buckets —
    // about 16 K elements while that needs at most 2048 bins (a tile of 16 K digits then still writes runs of >= 8 elements = one
    // sector), up to 32 K elements and 4096 bins beyond
    uint32_t lb = 0;
    while (lb < bits && (per_seg >> lb) > 16384) lb++;
    if (lb > 11) {
        lb = 11;
        while (lb < bits && (per_seg >> lb) > 32768) lb++;
    }
    if (bits > lb + 8) lb = bits - 8;                     // at most 8 fine bits (one byte per element, 256 counters)
    if (lb > 12) return false;                            // more than 4096 bins per segment: the tile runs would fall below a sector
    pl.nbins = 1u << lb;
    pl.F = bits - lb;
    pl.T = lb > 11 ? 32768 : 16384;
    const uint64_t expect = per_seg >> lb;
    pl.cap = (uint32_t)std::min<uint64_t>(40960, std::max<uint64_t>(4096, expect + expect / 4 + 1024));
    const size_t smem_digits = (size_t)pl.nseg * pl.nbins * 4, smem_part = ((size_t)3 * pl.nbins + 1) * 4 + (size_t)pl.T * 5;
    if (smem_digits > (size_t)smem_optin() - 1024 || smem_part > (size_t)smem_optin() - 1024) return false;
    return true;
}

size_t LaunchSort::coarse_words(const SortPlan &pl) { return (size_t)pl.nseg * pl.nbins + 1; }

template <class K> static void optin(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

int LaunchSort::run(cudaStream_t st, int scalar_field, int sm_count, const SortPlan &pl, const uint32_t *scalars, int is_mont, uint32_t *digits,
                    uint32_t *coarse_counts, uint32_t *coarse_off, uint32_t *cursor, uint32_t *part_pay, uint8_t *part_fine, uint32_t *entries, uint32_t *offsets,
                    cudaEvent_t ev_digits, cudaEvent_t ev_scan) {
    const uint32_t nh = pl.nseg * pl.nbins;
    const size_t smem_digits = (size_t)nh * 4;
    const unsigned grid_d = (unsigned)std::min<size_t>((pl.n + SORT_TPB - 1) / SORT_TPB, (size_t)sm_count * (smem_digits > 100 * 1024 ? 1 : 2));
    if (scalar_field == 0) {
        optin(k_sort_digits<FqP>, smem_digits);
        k_sort_digits<FqP><<<grid_d, SORT_TPB, smem_digits, st>>>(pl, scalars, is_mont, digits, coarse_counts);
    } else {
        optin(k_sort_digits<FrP>, smem_digits);
        k_sort_digits<FrP><<<grid_d, SORT_TPB, smem_digits, st>>>(pl, scalars, is_mont, digits, coarse_counts);
    }
    if (ev_digits) cudaEventRecord(ev_digits, st);
    const uint32_t G = pl.nseg * (pl.nbins << pl.F);
    k_sort_scan<<<1, SORT_TPB, 0, st>>>(nh, coarse_counts, coarse_off, cursor, offsets + G);
    if (ev_scan) cudaEventRecord(ev_scan, st);
    const size_t smem_part = ((size_t)3 * pl.nbins + 1) * 4 + (size_t)pl.T * 5;
    optin(k_sort_partition, smem_part);
    k_sort_partition<<<dim3((pl.n + pl.T - 1) / pl.T, pl.W), SORT_TPB, smem_part, st>>>(pl, digits, cursor, part_pay, part_fine);
    const size_t smem_b = (size_t)pl.cap * 4;
    optin(k_sort_buckets, smem_b);
    k_sort_buckets<<<std::min<uint32_t>(nh, 64u * (uint32_t)sm_count), SORT_TPB, smem_b, st>>>(pl, coarse_off, part_pay, part_fine, entries, offsets);
    return 4;
}

}  // namespace kgr
