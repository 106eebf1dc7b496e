// Shared device utilities for the B200 (sm_100a) ground-segmentation + clustering path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define LB_D __device__ __forceinline__
#define LB_HD __host__ __device__ __forceinline__

namespace lb
{

constexpr uint32_t kFullMask = 0xFFFFFFFFu;

// A batch is a set of independent frames laid end to end in every per-point array:
// frame f owns slots [off[f], off[f] + cnt[f]). off/cnt live in device memory because the obstacle
// counts of the clustering stage are produced on the device by the segmentation stage.
struct BatchView
{
    const uint32_t *off; // [F]   first slot of each frame
    const uint32_t *cnt; // [F]   live elements of each frame (<= capacity reserved at off)
    uint32_t frames;
};

// float -> uint32 whose unsigned order equals the float '<' order; -0.0 and +0.0 map to the same
// key because the reference comparators (`a.x < b.x`, segmentation.cpp:119-122) treat them as ties.
LB_HD uint32_t float_to_ordered(float f)
{
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union {
        float f;
        uint32_t u;
    } c;
    c.f = f;
    uint32_t u = c.u;
#endif
    if (u == 0x80000000u)
        u = 0u;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

LB_HD float ordered_to_float(uint32_t k)
{
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union {
        float f;
        uint32_t u;
    } c;
    c.u = u;
    return c.f;
#endif
}

LB_D uint32_t lane_id()
{
    return threadIdx.x & 31u;
}

LB_D uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

LB_D uint32_t warp_inclusive_scan(uint32_t v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t t = __shfl_up_sync(kFullMask, v, d);
        if (lane_id() >= static_cast<uint32_t>(d))
            v += t;
    }
    return v;
}

LB_D uint32_t warp_reduce_add(uint32_t v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v += __shfl_xor_sync(kFullMask, v, d);
    return v;
}

LB_D double warp_reduce_add(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v += __shfl_xor_sync(kFullMask, v, d);
    return v;
}

LB_D uint32_t warp_reduce_min(uint32_t v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v = min(v, __shfl_xor_sync(kFullMask, v, d));
    return v;
}

LB_D uint32_t warp_reduce_max(uint32_t v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v = max(v, __shfl_xor_sync(kFullMask, v, d));
    return v;
}

// Exclusive scan of one value per thread across a CTA of NT threads (NT multiple of 32, <= 1024).
// `ws` is NT/32 + 1 words of shared memory. Returns the exclusive prefix; *total gets the CTA sum.
// Contains two __syncthreads(); safe to call repeatedly with the same `ws`.
template <int NT> LB_D uint32_t block_exclusive_scan(uint32_t v, uint32_t *ws, uint32_t *total)
{
    constexpr int NW = NT / 32;
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t incl = warp_inclusive_scan(v);
    __syncthreads(); // protect ws from a previous call's readers
    if (lane == 31)
        ws[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = lane < NW ? ws[lane] : 0u;
        const uint32_t wi = warp_inclusive_scan(w);
        if (lane < NW)
            ws[lane] = wi - w;
        if (lane == 31)
            ws[NW] = wi;
    }
    __syncthreads();
    *total = ws[NW];
    return ws[warp] + incl - v;
}

// Sum of one 64-bit value per thread across a CTA of NT threads; every thread gets the total.
// `ws64` is NT/32 + 1 words of shared memory. Two __syncthreads(); safe to call repeatedly.
template <int NT> LB_D unsigned long long block_reduce_add64(unsigned long long v, unsigned long long *ws64)
{
    constexpr int NW = NT / 32;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v += __shfl_xor_sync(kFullMask, v, d);
    __syncthreads(); // protect ws64 from a previous call's readers
    if (lane_id() == 0)
        ws64[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0ull;
#pragma unroll
    for (int w = 0; w < NW; ++w)
        t += ws64[w];
    return t;
}

// Tiled scans: a frame is cut into tiles of one CTA each; pass 1 stores one packed 64-bit count per
// tile, pass 2 starts from the sum of the counts of the preceding tiles of its frame.
template <int NT>
LB_D unsigned long long tile_prefix64(const unsigned long long *__restrict__ tile_counts, uint32_t tile,
                                      unsigned long long *ws64)
{
    unsigned long long v = 0ull;
    for (uint32_t t = threadIdx.x; t < tile; t += NT)
        v += tile_counts[t];
    return block_reduce_add64<NT>(v, ws64);
}

// The reference build has no FMA contraction (x86-64 baseline, no -march; CMakeLists.txt has no
// arch flags), so squared distances and plane distances are formed with explicit round-to-nearest
// multiplies and adds. kdtree.hpp:145-163: d0 + (d1 + (d2 + 0)).
LB_D float dist_sqr_ref(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = __fsub_rn(ax, bx);
    const float dy = __fsub_rn(ay, by);
    const float dz = __fsub_rn(az, bz);
    const float d0 = __fmul_rn(dx, dx);
    const float d1 = __fmul_rn(dy, dy);
    const float d2 = __fmul_rn(dz, dz);
    return __fadd_rn(d0, __fadd_rn(d1, __fadd_rn(d2, 0.0f)));
}

} // namespace lb
