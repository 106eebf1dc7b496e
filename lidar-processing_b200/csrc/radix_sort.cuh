// Batched, stable LSD radix sort of (uint32 key, uint32 payload) pairs: every frame of a batch is
// sorted independently inside its own slot range. Hand-written for sm_100a: warp-synchronous
// multi-split ranking with __match_any_sync (no atomics on the ranking path), 8-bit digits.
//
// Used for (a) the x-order of Segmenter::form_planar_partitions (reference
// src/segmentation.cpp:116-122 — stable, ties by original index) and (b) grouping obstacle points by
// connected component in ascending index order for the clustering replay.
#pragma once

#include "common.cuh"

namespace lb
{

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems; // 4096 keys per CTA
constexpr int kRsRadix = 256;

// Per-tile digit histogram. grid = (max_tiles, frames). Tiles past the end of a frame write zeros
// so the scan below can run over a fixed-size table.
__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const uint32_t *__restrict__ keys, BatchView bv, uint32_t shift, uint32_t max_tiles,
               uint32_t *__restrict__ tile_hist)
{
    __shared__ uint32_t cnt[kRsRadix];
    const uint32_t f = blockIdx.y;
    const uint32_t tile = blockIdx.x;
    const uint32_t n = bv.cnt[f];
    const uint32_t off = bv.off[f];
    cnt[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t base = tile * kRsTile;
    if (base < n)
    {
#pragma unroll 4
        for (int k = 0; k < kRsItems; ++k)
        {
            const uint32_t idx = base + k * kRsThreads + threadIdx.x;
            const bool valid = idx < n;
            const uint32_t digit = valid ? ((keys[off + idx] >> shift) & 0xFFu) : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(kFullMask, digit);
            if (valid && (peers & lanemask_lt()) == 0u)
                atomicAdd(&cnt[digit], static_cast<uint32_t>(__popc(peers)));
        }
    }
    __syncthreads();
    tile_hist[(static_cast<size_t>(f) * kRsRadix + threadIdx.x) * max_tiles + tile] = cnt[threadIdx.x];
}

// Exclusive scan of each frame's (digit-major, tile-minor) table. grid = (frames), 1024 threads.
__global__ void __launch_bounds__(1024) rs_scan_kernel(uint32_t *__restrict__ tile_hist, uint32_t entries_per_frame)
{
    constexpr int kPer = 8;
    __shared__ uint32_t ws[33];
    uint32_t *tab = tile_hist + static_cast<size_t>(blockIdx.x) * entries_per_frame;
    uint32_t carry = 0u;
    for (uint32_t base = 0; base < entries_per_frame; base += 1024 * kPer)
    {
        uint32_t v[kPer];
        uint32_t sum = 0u;
        const uint32_t first = base + threadIdx.x * kPer;
#pragma unroll
        for (int k = 0; k < kPer; ++k)
        {
            v[k] = (first + k < entries_per_frame) ? tab[first + k] : 0u;
            sum += v[k];
        }
        uint32_t total;
        uint32_t run = carry + block_exclusive_scan<1024>(sum, ws, &total);
#pragma unroll
        for (int k = 0; k < kPer; ++k)
        {
            if (first + k < entries_per_frame)
                tab[first + k] = run;
            run += v[k];
        }
        carry += total;
    }
}

// Stable scatter of one tile. grid = (max_tiles, frames). Warp w owns the contiguous sub-range
// [tile_base + w*512, +512); inside it, iteration k / lane l maps to element k*32 + l, so
// (warp, k, lane) order equals memory order and equal digits keep their relative order.
__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, BatchView bv, uint32_t shift,
                  uint32_t max_tiles, const uint32_t *__restrict__ tile_hist)
{
    constexpr int kWarps = kRsThreads / 32;
    constexpr int kPerWarp = kRsTile / kWarps; // 512
    __shared__ uint32_t warp_cnt[kWarps][kRsRadix];

    const uint32_t f = blockIdx.y;
    const uint32_t tile = blockIdx.x;
    const uint32_t n = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t base = tile * kRsTile;
    if (base >= n)
        return;

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    for (int d = lane; d < kRsRadix; d += 32)
        warp_cnt[warp][d] = 0u;
    __syncwarp();

    uint32_t key[kRsItems];
    uint32_t val[kRsItems];
    uint16_t rank[kRsItems];
    const uint32_t wbase = base + warp * kPerWarp;
#pragma unroll
    for (int k = 0; k < kRsItems; ++k)
    {
        const uint32_t idx = wbase + k * 32 + lane;
        const bool valid = idx < n;
        key[k] = valid ? keys_in[off + idx] : 0u;
        val[k] = valid ? vals_in[off + idx] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kRsItems; ++k)
    {
        const uint32_t idx = wbase + k * 32 + lane;
        const bool valid = idx < n;
        const uint32_t digit = valid ? ((key[k] >> shift) & 0xFFu) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(kFullMask, digit);
        uint32_t prev = 0u;
        if (valid)
            prev = warp_cnt[warp][digit];
        __syncwarp();
        if (valid && (peers & lanemask_lt()) == 0u)
            warp_cnt[warp][digit] = prev + static_cast<uint32_t>(__popc(peers));
        __syncwarp();
        rank[k] = static_cast<uint16_t>(prev + static_cast<uint32_t>(__popc(peers & lanemask_lt())));
    }
    __syncthreads();
    {
        // thread d turns the per-warp counts of digit d into global destinations
        const uint32_t d = threadIdx.x;
        uint32_t run = tile_hist[(static_cast<size_t>(f) * kRsRadix + d) * max_tiles + tile];
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
        {
            const uint32_t c = warp_cnt[w][d];
            warp_cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kRsItems; ++k)
    {
        const uint32_t idx = wbase + k * 32 + lane;
        if (idx < n)
        {
            const uint32_t digit = (key[k] >> shift) & 0xFFu;
            const uint32_t dst = off + warp_cnt[warp][digit] + rank[k];
            keys_out[dst] = key[k];
            vals_out[dst] = val[k];
        }
    }
}

struct RadixSortScratch
{
    uint32_t *tile_hist; // frames * 256 * max_tiles words
};

inline size_t radix_sort_scratch_words(uint32_t frames, uint32_t max_n)
{
    const uint32_t max_tiles = (max_n + kRsTile - 1) / kRsTile;
    return static_cast<size_t>(frames) * kRsRadix * (max_tiles ? max_tiles : 1u);
}

// Sorts by the low `bits` bits of the key (rounded up to a multiple of 8). Returns the number of
// passes executed: after an odd number the result is in (keys_b, vals_b), else in (keys_a, vals_a).
inline int radix_sort_pairs(cudaStream_t stream, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b,
                            BatchView bv, uint32_t max_n, uint32_t bits, RadixSortScratch scratch, int *launches)
{
    if (max_n == 0u || bv.frames == 0u)
        return 0;
    const uint32_t max_tiles = (max_n + kRsTile - 1) / kRsTile;
    const int passes = static_cast<int>((bits + 7u) / 8u);
    const dim3 grid(max_tiles, bv.frames);
    uint32_t *kin = keys_a, *vin = vals_a, *kout = keys_b, *vout = vals_b;
    for (int p = 0; p < passes; ++p)
    {
        const uint32_t shift = static_cast<uint32_t>(p) * 8u;
        rs_hist_kernel<<<grid, kRsThreads, 0, stream>>>(kin, bv, shift, max_tiles, scratch.tile_hist);
        rs_scan_kernel<<<bv.frames, 1024, 0, stream>>>(scratch.tile_hist, kRsRadix * max_tiles);
        rs_scatter_kernel<<<grid, kRsThreads, 0, stream>>>(kin, vin, kout, vout, bv, shift, max_tiles,
                                                            scratch.tile_hist);
        if (launches)
            *launches += 3;
        uint32_t *t = kin;
        kin = kout;
        kout = t;
        t = vin;
        vin = vout;
        vout = t;
    }
    return passes;
}

} // namespace lb
