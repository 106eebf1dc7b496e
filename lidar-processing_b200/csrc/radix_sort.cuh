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
constexpr int kRsScatterThreads = 512; // the scatter keeps kRsTile / 512 = 8 pairs per thread in registers: <= 40 registers,
constexpr int kRsScatterItems = kRsTile / kRsScatterThreads; // 48 warps per SM hide the load latency it is bound by

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
// [tile_base + w*256, +256); inside it, iteration k / lane l maps to element k*32 + l, so
// (warp, k, lane) order equals memory order and equal digits keep their relative order.
// The tile is ordered by digit in shared memory first and leaves in runs of equal digit: a direct scatter touches ~29
// sectors per 32-lane store (ncu: L2 at 71 % of its sector throughput, DRAM at 16 %), the staged one 4-8.
__global__ void __launch_bounds__(kRsScatterThreads, 3)
rs_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, BatchView bv, uint32_t shift,
                  uint32_t max_tiles, const uint32_t *__restrict__ tile_hist)
{
    constexpr int kWarps = kRsScatterThreads / 32;
    constexpr int kPerWarp = kRsTile / kWarps; // 256
    __shared__ uint32_t stage_key[kRsTile];
    __shared__ uint32_t stage_val[kRsTile];
    __shared__ uint32_t gbase[kRsRadix];      // first destination of the tile's run of every digit (frame-relative)
    __shared__ uint32_t tile_excl[kRsRadix];  // first staging slot of every digit
    __shared__ uint16_t warp_cnt[kWarps][kRsRadix]; // per warp: count, then first staging slot relative to tile_excl
    __shared__ uint32_t ws[kRsRadix / 32];

    const uint32_t f = blockIdx.y;
    const uint32_t tile = blockIdx.x;
    const uint32_t n = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t base = tile * kRsTile;
    if (base >= n)
        return;

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const uint32_t lt = lanemask_lt();
    for (int d = lane; d < kRsRadix; d += 32)
        warp_cnt[warp][d] = 0u;
    __syncwarp();

    uint32_t key[kRsScatterItems];
    uint32_t val[kRsScatterItems];
    uint16_t rank[kRsScatterItems];
    const uint32_t wbase = base + warp * kPerWarp;
#pragma unroll
    for (int k = 0; k < kRsScatterItems; ++k)
    {
        const uint32_t idx = wbase + k * 32 + lane;
        const bool valid = idx < n;
        key[k] = valid ? keys_in[off + idx] : 0u;
        val[k] = valid ? vals_in[off + idx] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kRsScatterItems; ++k)
    {
        const uint32_t idx = wbase + k * 32 + lane;
        const bool valid = idx < n;
        const uint32_t digit = valid ? ((key[k] >> shift) & 0xFFu) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(kFullMask, digit);
        uint32_t prev = 0u;
        if (valid)
            prev = warp_cnt[warp][digit];
        __syncwarp();
        if (valid && (peers & lt) == 0u)
            warp_cnt[warp][digit] = static_cast<uint16_t>(prev + __popc(peers));
        __syncwarp();
        rank[k] = static_cast<uint16_t>(prev + __popc(peers & lt));
    }
    __syncthreads();
    {
        // thread d: counts of the warps for digit d -> first slot of (warp, d); the tile's count of d is scanned over the digits
        const uint32_t d = threadIdx.x;
        uint32_t c = 0u;
        if (d < kRsRadix)
        {
            gbase[d] = tile_hist[(static_cast<size_t>(f) * kRsRadix + d) * max_tiles + tile];
#pragma unroll
            for (int w = 0; w < kWarps; ++w)
            {
                const uint32_t x = warp_cnt[w][d];
                warp_cnt[w][d] = static_cast<uint16_t>(c);
                c += x;
            }
        }
        const uint32_t incl = warp_inclusive_scan(c);
        if (lane == 31 && warp < kRsRadix / 32)
            ws[warp] = incl;
        __syncthreads();
        if (d < kRsRadix)
        {
            uint32_t before = 0u;
            for (uint32_t v = 0; v < warp; ++v)
                before += ws[v];
            tile_excl[d] = before + incl - c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kRsScatterItems; ++k)
    {
        if (wbase + k * 32 + lane < n)
        {
            const uint32_t digit = (key[k] >> shift) & 0xFFu;
            const uint32_t slot = tile_excl[digit] + warp_cnt[warp][digit] + rank[k];
            stage_key[slot] = key[k];
            stage_val[slot] = val[k];
        }
    }
    __syncthreads();
    const uint32_t tn = min(static_cast<uint32_t>(kRsTile), n - base);
    for (uint32_t j = threadIdx.x; j < tn; j += kRsScatterThreads)
    {
        const uint32_t k2 = stage_key[j];
        const uint32_t digit = (k2 >> shift) & 0xFFu;
        const uint32_t dst = off + gbase[digit] + (j - tile_excl[digit]);
        keys_out[dst] = k2;
        vals_out[dst] = stage_val[j];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Frame-resident variant for large batches: ONE CTA sorts one frame through ALL passes (a persistent grid walks the
// frames), so a batch costs one launch instead of three per pass, every pass after the first reads what the same SM
// just wrote (L2-resident: a frame is ~1 MB of pairs), and the scatter is staged through shared memory - a tile of
// 4096 pairs is ordered by digit there and leaves in runs of equal digit, i.e. in whole 32-byte sectors instead of one
// 4-byte store per sector (the direct scatter above spends its time on that write amplification).
//   * the digit histograms of all passes are taken in one read of the keys before the first pass,
//   * tiles are processed in memory order and (warp, iteration, lane) order equals memory order inside a tile, so equal
//     digits keep their relative order: the sort is stable like the per-tile version,
//   * a pass whose digit is the same for every key of the frame is a straight copy.
// Used for the component sort (2-3 passes over ~50 k pairs per frame: 0.30 ms per 154 frames against 0.6 ms). For the
// 4-pass x sort of the whole cloud one CTA per frame is too little parallelism (16 warps per SM): measured slower.
constexpr int kFsThreads = 512;
constexpr int kFsItems = 8;
constexpr int kFsWarps = kFsThreads / 32;
constexpr int kFsTile = kFsThreads * kFsItems; // 4096 pairs
constexpr int kFsMaxPasses = 4;

struct __align__(16) FrameSortSmem
{
    uint32_t stage_key[kFsTile];
    uint32_t stage_val[kFsTile];
    uint32_t hist[kFsMaxPasses][kRsRadix]; // frame-wide digit counts per pass
    uint32_t bin_base[kRsRadix];           // next free destination of every digit (frame-relative)
    uint32_t tile_excl[kRsRadix];          // first staging slot of every digit in the current tile
    uint16_t warp_cnt[kFsWarps][kRsRadix]; // per warp: count, then first staging slot of (warp, digit) relative to tile_excl
    uint32_t ws[kFsWarps + 1];
};

// grid = min(frames, resident CTAs); dynamic shared memory = sizeof(FrameSortSmem).
// Pass 0 reads (keys_a, vals_a); pass p writes buffer b / a alternately.
__global__ void __launch_bounds__(kFsThreads, 2)
rs_frame_sort_kernel(uint32_t *__restrict__ keys_a, uint32_t *__restrict__ vals_a, uint32_t *__restrict__ keys_b,
                     uint32_t *__restrict__ vals_b, BatchView bv, int passes)
{
    extern __shared__ __align__(16) unsigned char fs_raw[];
    FrameSortSmem &sm = *reinterpret_cast<FrameSortSmem *>(fs_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    for (uint32_t f = blockIdx.x; f < bv.frames; f += gridDim.x)
    {
        const uint32_t n = bv.cnt[f];
        const uint32_t off = bv.off[f];
        if (n == 0u)
            continue;
        __syncthreads();
        for (uint32_t i = tid; i < kFsMaxPasses * kRsRadix; i += kFsThreads)
            (&sm.hist[0][0])[i] = 0u;
        __syncthreads();
        // ---- digit histograms of every pass in one read
        for (uint32_t base = 0; base < n; base += kFsThreads)
        {
            const uint32_t i = base + tid;
            const bool valid = i < n;
            const uint32_t key = valid ? keys_a[off + i] : 0u;
            for (int p = 0; p < passes; ++p)
            {
                const uint32_t digit = valid ? ((key >> (8 * p)) & 0xFFu) : 0xFFFFFFFFu;
                const uint32_t peers = __match_any_sync(kFullMask, digit);
                if (valid && (peers & lt) == 0u)
                    atomicAdd(&sm.hist[p][digit], static_cast<uint32_t>(__popc(peers)));
            }
        }
        __syncthreads();
        for (int p = 0; p < passes; ++p)
        {
            const uint32_t shift = 8u * static_cast<uint32_t>(p);
            const uint32_t *kin = (p & 1) ? keys_b + off : keys_a + off;
            const uint32_t *vin = (p & 1) ? vals_b + off : vals_a + off;
            uint32_t *kout = (p & 1) ? keys_a + off : keys_b + off;
            uint32_t *vout = (p & 1) ? vals_a + off : vals_b + off;
            // ---- bin_base = exclusive scan of the pass's histogram; a constant digit makes the pass a copy
            bool constant;
            {
                const uint32_t c = tid < kRsRadix ? sm.hist[p][tid] : 0u;
                const uint32_t incl = warp_inclusive_scan(c);
                if (lane == 31)
                    sm.ws[warp] = incl;
                const int any_full = __syncthreads_or(c == n ? 1 : 0);
                uint32_t before = 0u;
                for (uint32_t v = 0; v < warp && v < kRsRadix / 32u; ++v)
                    before += sm.ws[v];
                if (tid < kRsRadix)
                    sm.bin_base[tid] = before + incl - c;
                constant = any_full != 0;
                __syncthreads();
            }
            if (constant)
            {
                for (uint32_t i = tid; i < n; i += kFsThreads)
                {
                    kout[i] = kin[i];
                    vout[i] = vin[i];
                }
                __syncthreads();
                continue;
            }
            for (uint32_t tbase = 0; tbase < n; tbase += kFsTile)
            {
                // ---- load + rank inside the warp (warp w owns kFsItems * 32 consecutive pairs of the tile, in memory order)
                for (uint32_t d = lane; d < kRsRadix; d += 32u)
                    sm.warp_cnt[warp][d] = 0u;
                __syncwarp();
                uint32_t key[kFsItems], val[kFsItems];
                uint16_t rank[kFsItems];
                const uint32_t wbase = tbase + warp * (kFsItems * 32u);
#pragma unroll
                for (int k = 0; k < kFsItems; ++k)
                {
                    const uint32_t i = wbase + k * 32 + lane;
                    const bool valid = i < n;
                    key[k] = valid ? kin[i] : 0u;
                    val[k] = valid ? vin[i] : 0u;
                }
#pragma unroll
                for (int k = 0; k < kFsItems; ++k)
                {
                    const bool valid = wbase + k * 32 + lane < n;
                    const uint32_t digit = valid ? ((key[k] >> shift) & 0xFFu) : 0xFFFFFFFFu;
                    const uint32_t peers = __match_any_sync(kFullMask, digit);
                    uint32_t prev = 0u;
                    if (valid)
                        prev = sm.warp_cnt[warp][digit];
                    __syncwarp();
                    if (valid && (peers & lt) == 0u)
                        sm.warp_cnt[warp][digit] = static_cast<uint16_t>(prev + __popc(peers));
                    __syncwarp();
                    rank[k] = static_cast<uint16_t>(prev + __popc(peers & lt));
                }
                __syncthreads();
                // ---- per digit: counts of the warps -> first slot of (warp, digit); tile_excl = scan over the digits
                {
                    uint32_t c = 0u;
                    if (tid < kRsRadix)
                    {
#pragma unroll
                        for (int w = 0; w < kFsWarps; ++w)
                        {
                            const uint32_t x = sm.warp_cnt[w][tid];
                            sm.warp_cnt[w][tid] = static_cast<uint16_t>(c);
                            c += x;
                        }
                    }
                    const uint32_t incl = warp_inclusive_scan(c);
                    if (lane == 31)
                        sm.ws[warp] = incl;
                    __syncthreads();
                    uint32_t before = 0u;
                    for (uint32_t v = 0; v < warp && v < kRsRadix / 32u; ++v)
                        before += sm.ws[v];
                    if (tid < kRsRadix)
                        sm.tile_excl[tid] = before + incl - c;
                    __syncthreads();
                }
                // ---- stage the tile in digit order
#pragma unroll
                for (int k = 0; k < kFsItems; ++k)
                {
                    if (wbase + k * 32 + lane < n)
                    {
                        const uint32_t digit = (key[k] >> shift) & 0xFFu;
                        const uint32_t slot = sm.tile_excl[digit] + sm.warp_cnt[warp][digit] + rank[k];
                        sm.stage_key[slot] = key[k];
                        sm.stage_val[slot] = val[k];
                    }
                }
                __syncthreads();
                // ---- runs of equal digit leave for consecutive destinations
                const uint32_t tn = min(static_cast<uint32_t>(kFsTile), n - tbase);
                for (uint32_t j = tid; j < tn; j += kFsThreads)
                {
                    const uint32_t k2 = sm.stage_key[j], v2 = sm.stage_val[j];
                    const uint32_t digit = (k2 >> shift) & 0xFFu;
                    const uint32_t dst = sm.bin_base[digit] + (j - sm.tile_excl[digit]);
                    kout[dst] = k2;
                    vout[dst] = v2;
                }
                __syncthreads();
                if (tid < kRsRadix)
                {
                    const uint32_t next = tid + 1u < kRsRadix ? sm.tile_excl[tid + 1u] : tn;
                    sm.bin_base[tid] += next - sm.tile_excl[tid];
                }
                __syncthreads();
            }
        }
    }
}

// Batch-size switch: the frame-resident kernel needs enough frames to occupy the machine (one CTA per frame).
constexpr uint32_t kFsMinFrames = 32u;

inline cudaError_t rs_frame_sort_launch(cudaStream_t stream, uint32_t sm_count, uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b,
                                        uint32_t *vals_b, BatchView bv, int passes)
{
    static bool attr_done = false; // (a per-device function attribute; contexts of one process share the device)
    if (!attr_done)
    {
        const cudaError_t e = cudaFuncSetAttribute(rs_frame_sort_kernel,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(FrameSortSmem)));
        if (e != cudaSuccess)
            return e;
        attr_done = true;
    }
    const uint32_t grid = bv.frames < 2u * sm_count ? bv.frames : 2u * sm_count;
    rs_frame_sort_kernel<<<grid, kFsThreads, sizeof(FrameSortSmem), stream>>>(keys_a, vals_a, keys_b, vals_b, bv, passes);
    return cudaSuccess;
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
        rs_scatter_kernel<<<grid, kRsScatterThreads, 0, stream>>>(kin, vin, kout, vout, bv, shift, max_tiles,
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
