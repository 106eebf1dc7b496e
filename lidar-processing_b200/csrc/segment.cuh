// Ground segmentation kernels (sm_100a). Drop-in for lidar_processing::Segmenter::segment
// (reference src/segmentation.cpp:311-345 and the helpers it calls, 62-309).
//
// Stages for a batch of frames (every kernel covers all frames of the batch):
//   seg_keys     float4 point load -> order-preserving x keys + iota          (segmentation.cpp:116-117)
//   radix sort   stable x order, ties by original index                       (segmentation.cpp:119-122)
//   seg_gather   x-sorted float4 copy {x, y, z, original index}               (segmentation.cpp:137-144)
//   seg_fit      ONE CTA per (frame, planar partition): exact radix-select of the
//                lowest-point representatives, their ascending sequential float sum,
//                seed cut, then `iterations` x { moments -> 3x3 covariance ->
//                Jacobi SVD -> signed point-to-plane classify }                (segmentation.cpp:151-309)
//   seg_compact  stable compaction to ground / obstacle lists + label scatter  (segmentation.cpp:331-343)
#pragma once

#include "common.cuh"
#include "jacobi3.h"

namespace lb
{

constexpr uint32_t kSegUnknown = 0u;
constexpr uint32_t kSegGround = 1u;
constexpr uint32_t kSegObstacle = 2u;

constexpr int kFitThreads = 1024;
constexpr int kFitUnroll = 4; // independent loads in flight per thread in every pass of seg_fit
constexpr uint32_t kMaxLpr = 8192u; // lower point representatives sorted in shared memory; more of them go through global memory

struct SegParams
{
    float sensor_height_m;
    float orthogonal_distance_threshold;
    float initial_seed_threshold;
    uint32_t iterations;
    uint32_t partitions;
    uint32_t lpr;
};

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
seg_keys_kernel(const float4 *__restrict__ pts, BatchView bv, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const uint32_t f = blockIdx.y;
    const uint32_t n = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 p = __ldg(&pts[off + i]);
        keys[off + i] = float_to_ordered(p.x);
        vals[off + i] = i;
    }
}

__global__ void __launch_bounds__(256)
seg_gather_kernel(const float4 *__restrict__ pts, const uint32_t *__restrict__ sorted_idx, BatchView bv,
                  float4 *__restrict__ spts, uint32_t *__restrict__ zkeys)
{
    const uint32_t f = blockIdx.y;
    const uint32_t n = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    {
        const uint32_t idx = sorted_idx[off + r];
        const float4 p = __ldg(&pts[off + idx]);
        spts[off + r] = make_float4(p.x, p.y, p.z, __uint_as_float(idx));
        zkeys[off + r] = float_to_ordered(p.z); // 4-byte keys for the selection passes of seg_fit (coalesced)
    }
}

// ---------------------------------------------------------------------------------------------
// seg_fit: shared-memory layout (dynamic)
struct FitSmem
{
    uint32_t hist[4096];
    uint32_t buf[kMaxLpr];
    double red[32][11];
    uint32_t ws[40];
    // broadcast slots
    uint32_t u[16];
    float fl[16];
};

// finds the bin holding the `target`-th (1-based) element of a histogram; every thread gets the
// bin and the number of elements in lower bins. nbins must be a multiple of 4 and <= 4096.
LB_D void fit_find_bin(FitSmem &sm, uint32_t nbins, uint32_t target, uint32_t *bin_out, uint32_t *before_out)
{
    const uint32_t t = threadIdx.x;
    uint32_t h[4] = {0u, 0u, 0u, 0u};
    uint32_t sum = 0u;
    if (t * 4 < nbins)
    {
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            h[k] = sm.hist[t * 4 + k];
            sum += h[k];
        }
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan<kFitThreads>(sum, sm.ws, &total);
    if (t * 4 < nbins)
    {
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            if (h[k] != 0u && run < target && target <= run + h[k])
            {
                sm.u[0] = t * 4 + k;
                sm.u[1] = run;
            }
            run += h[k];
        }
    }
    __syncthreads();
    *bin_out = sm.u[0];
    *before_out = sm.u[1];
    __syncthreads();
}

LB_D void fit_hist_add(FitSmem &sm, bool valid, uint32_t bin)
{
    const uint32_t peers = __match_any_sync(kFullMask, valid ? bin : 0xFFFFFFFFu);
    if (valid && (peers & lanemask_lt()) == 0u)
        atomicAdd(&sm.hist[bin], static_cast<uint32_t>(__popc(peers)));
}

// block-wide sum of 10 moment accumulators + a count; result valid in every thread via sm.red[0]
LB_D void fit_reduce_moments(FitSmem &sm, double acc[10], uint32_t count, double out[10], uint32_t *count_out)
{
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    double c = static_cast<double>(count);
#pragma unroll
    for (int k = 0; k < 10; ++k)
        acc[k] = warp_reduce_add(acc[k]);
    c = warp_reduce_add(c);
    __syncthreads();
    if (lane == 0)
    {
#pragma unroll
        for (int k = 0; k < 10; ++k)
            sm.red[warp][k] = acc[k];
        sm.red[warp][10] = c;
    }
    __syncthreads();
    if (warp == 0)
    {
#pragma unroll
        for (int k = 0; k < 11; ++k)
        {
            double v = sm.red[lane][k];
            v = warp_reduce_add(v);
            if (lane == 0)
                sm.red[0][k] = v;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 10; ++k)
        out[k] = sm.red[0][k];
    *count_out = static_cast<uint32_t>(sm.red[0][10]);
    __syncthreads();
}

struct PlaneF
{
    float a, b, c, d, thr;
};

// Segmenter::estimate_plane_coefficients (segmentation.cpp:62-102) from raw moments taken about
// the shift point (sx, sy, sz). Moments are accumulated in double so that the covariance is the
// correctly-rounded one; the reference's float GEMM summation order is not knowable.
LB_D bool fit_plane_from_moments(const double m[10], uint32_t n, double sx, double sy, double sz, float odt,
                                 PlaneF *pl)
{
    if (n < 3u)
        return false;
    const double dn = static_cast<double>(n);
    const double mx = m[0] / dn, my = m[1] / dn, mz = m[2] / dn;
    const double dd = static_cast<double>(n - 1u);
    float cov[9];
    cov[0] = static_cast<float>((m[3] - dn * mx * mx) / dd);
    cov[1] = static_cast<float>((m[4] - dn * mx * my) / dd);
    cov[2] = static_cast<float>((m[5] - dn * mx * mz) / dd);
    cov[4] = static_cast<float>((m[6] - dn * my * my) / dd);
    cov[5] = static_cast<float>((m[7] - dn * my * mz) / dd);
    cov[8] = static_cast<float>((m[8] - dn * mz * mz) / dd);
    cov[3] = cov[1];
    cov[6] = cov[2];
    cov[7] = cov[5];
    float v[9], sv[3];
    if (!jacobi_svd3(cov, v, sv))
        return false;
    const float cx = static_cast<float>(sx + mx);
    const float cy = static_cast<float>(sy + my);
    const float cz = static_cast<float>(sz + mz);
    pl->a = v[2];
    pl->b = v[5];
    pl->c = v[8];
    pl->d = __fadd_rn(__fadd_rn(__fmul_rn(pl->a, cx), __fmul_rn(pl->b, cy)), __fmul_rn(pl->c, cz));
    const float nn = __fadd_rn(__fadd_rn(__fmul_rn(pl->a, pl->a), __fmul_rn(pl->b, pl->b)), __fmul_rn(pl->c, pl->c));
    pl->thr = __fmul_rn(odt, __fsqrt_rn(nn)); // segmentation.cpp:293
    return true;
}

// signed point-to-plane test (segmentation.cpp:290-299): (x*a + y*b) + z*c - d < 0.3*|n|
LB_D bool fit_is_ground(const float4 &p, const PlaneF &pl)
{
    const float dist =
        __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(p.x, pl.a), __fmul_rn(p.y, pl.b)), __fmul_rn(p.z, pl.c)), pl.d);
    return dist < pl.thr;
}

// grid = (partitions, frames), kFitThreads threads, dynamic smem = sizeof(FitSmem)
// status: 0 ok, 1 "<3 points" (points stay UNKNOWN), 2 "Failed ground segmentation" (all OBSTACLE)
__global__ void __launch_bounds__(kFitThreads, 1)
seg_fit_kernel(const float4 *__restrict__ spts, const uint32_t *__restrict__ zkeys, BatchView bv, SegParams prm,
               uint8_t *__restrict__ flags,
               float *__restrict__ planes_out, int32_t *__restrict__ status_out, uint32_t *__restrict__ lpr_spill)
{
    extern __shared__ __align__(16) unsigned char fit_smem_raw[];
    FitSmem &sm = *reinterpret_cast<FitSmem *>(fit_smem_raw);

    const uint32_t f = blockIdx.y;
    const uint32_t s = blockIdx.x;
    const uint32_t n = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t P = prm.partitions;
    const uint32_t per = n / P; // segmentation.cpp:124
    const uint32_t lo = s * per;
    const uint32_t hi = lo + per;
    const uint32_t tid = threadIdx.x;
    const float4 *seg = spts + off;
    const uint32_t *zk = zkeys + off;
    uint8_t *fl = flags + off;
    float *planes = planes_out + (static_cast<size_t>(f) * P + s) * prm.iterations * 4;
    int32_t *status = status_out + static_cast<size_t>(f) * P + s;

    for (uint32_t i = tid; i < prm.iterations * 4; i += kFitThreads)
        planes[i] = __int_as_float(0x7FC00000);

    // points past partitions*per belong to no partition and stay UNKNOWN (segmentation.cpp:124-148)
    if (s == P - 1u)
        for (uint32_t i = hi + tid; i < n; i += kFitThreads)
            fl[i] = 0u;

    if (per < 3u) // segmentation.cpp:225-229
    {
        for (uint32_t i = lo + tid; i < hi; i += kFitThreads)
            fl[i] = 0u;
        if (tid == 0)
            *status = 1;
        return;
    }

    // ---- extract_initial_seeds (segmentation.cpp:151-217) ----
    const uint32_t kmin = float_to_ordered(__fmul_rn(-1.5f, prm.sensor_height_m));

    // level 1: histogram of key bits [31:20] over ALL points + count above z_min + max key
    for (uint32_t i = tid; i < 4096u; i += kFitThreads)
        sm.hist[i] = 0u;
    __syncthreads();
    uint32_t cnt_above = 0u, kmax = 0u;
    for (uint32_t base = lo; base < hi; base += kFitThreads * kFitUnroll)
    {
        uint32_t k[kFitUnroll];
        bool valid[kFitUnroll];
#pragma unroll
        for (int h = 0; h < kFitUnroll; ++h)
        {
            const uint32_t i = base + h * kFitThreads + tid;
            valid[h] = i < hi;
            k[h] = valid[h] ? __ldg(&zk[i]) : 0u;
        }
#pragma unroll
        for (int h = 0; h < kFitUnroll; ++h)
        {
            if (valid[h])
            {
                cnt_above += (k[h] > kmin) ? 1u : 0u;
                kmax = max(kmax, k[h]);
            }
            fit_hist_add(sm, valid[h], k[h] >> 20);
        }
    }
    cnt_above = warp_reduce_add(cnt_above);
    kmax = warp_reduce_max(kmax);
    if (lane_id() == 0)
    {
        sm.ws[tid >> 5] = cnt_above; // reuse as scratch: 32 words
    }
    __syncthreads();
    if (tid < 32)
    {
        uint32_t v = warp_reduce_add(sm.ws[tid]);
        if (tid == 0)
            sm.u[2] = v;
    }
    __syncthreads();
    cnt_above = sm.u[2];
    __syncthreads();
    if (lane_id() == 0)
        sm.ws[tid >> 5] = kmax;
    __syncthreads();
    if (tid < 32)
    {
        uint32_t v = warp_reduce_max(sm.ws[tid]);
        if (tid == 0)
            sm.u[3] = v;
    }
    __syncthreads();
    kmax = sm.u[3];
    __syncthreads();

    // If no point lies above z_min nothing is erased (cutoff index stays 0, segmentation.cpp:171-182)
    const bool use_kmin = cnt_above != 0u;
    const uint32_t n_inc = use_kmin ? cnt_above : per;
    const uint32_t n_lpr = min(n_inc, prm.lpr);
    if (use_kmin)
    {
        // remove the excluded keys (k <= kmin) from the level-1 histogram
        const uint32_t bk = kmin >> 20;
        if (tid == 0)
        {
            uint32_t below = 0u;
            for (uint32_t b = 0; b < bk; ++b)
                below += sm.hist[b];
            sm.u[4] = below;
        }
        __syncthreads();
        const uint32_t below = sm.u[4];
        __syncthreads();
        for (uint32_t b = tid; b < bk; b += kFitThreads)
            sm.hist[b] = 0u;
        if (tid == 0)
            sm.hist[bk] -= (per - cnt_above) - below;
        __syncthreads();
    }
    uint32_t b1, before1;
    fit_find_bin(sm, 4096u, n_lpr, &b1, &before1);
    uint32_t remaining = n_lpr - before1;

    // level 2: bits [19:8] among included keys with top-12 == b1
    for (uint32_t i = tid; i < 4096u; i += kFitThreads)
        sm.hist[i] = 0u;
    __syncthreads();
    for (uint32_t base = lo; base < hi; base += kFitThreads * kFitUnroll)
    {
        uint32_t k[kFitUnroll];
        bool valid[kFitUnroll];
#pragma unroll
        for (int h = 0; h < kFitUnroll; ++h)
        {
            const uint32_t i = base + h * kFitThreads + tid;
            valid[h] = i < hi;
            k[h] = valid[h] ? __ldg(&zk[i]) : 0u;
        }
#pragma unroll
        for (int h = 0; h < kFitUnroll; ++h)
        {
            const bool v = valid[h] && (!use_kmin || k[h] > kmin) && (k[h] >> 20) == b1;
            if (__any_sync(kFullMask, v)) // most warps hold no key of the selected bin
                fit_hist_add(sm, v, (k[h] >> 8) & 0xFFFu);
        }
    }
    __syncthreads();
    uint32_t b2, before2;
    fit_find_bin(sm, 4096u, remaining, &b2, &before2);
    remaining -= before2;

    // level 3: bits [7:0] among included keys with top-24 == (b1,b2)
    for (uint32_t i = tid; i < 4096u; i += kFitThreads)
        sm.hist[i] = 0u;
    __syncthreads();
    const uint32_t prefix24 = (b1 << 12) | b2;
    for (uint32_t base = lo; base < hi; base += kFitThreads * kFitUnroll)
    {
        uint32_t k[kFitUnroll];
        bool valid[kFitUnroll];
#pragma unroll
        for (int h = 0; h < kFitUnroll; ++h)
        {
            const uint32_t i = base + h * kFitThreads + tid;
            valid[h] = i < hi;
            k[h] = valid[h] ? __ldg(&zk[i]) : 0u;
        }
#pragma unroll
        for (int h = 0; h < kFitUnroll; ++h)
        {
            const bool v = valid[h] && (!use_kmin || k[h] > kmin) && (k[h] >> 8) == prefix24;
            if (__any_sync(kFullMask, v))
                fit_hist_add(sm, v, k[h] & 0xFFu);
        }
    }
    __syncthreads();
    uint32_t b3, before3;
    fit_find_bin(sm, 256u, remaining, &b3, &before3);
    remaining -= before3; // copies of the threshold value T that belong to the lowest n_lpr (>= 1)
    const uint32_t kT = (prefix24 << 8) | b3;
    const uint32_t c_less = n_lpr - remaining; // included keys strictly below T  (<= lpr - 1)

    if (c_less > kMaxLpr)
    {
        // More lower point representatives than the shared-memory sort holds (dense sensors, merged clouds): the keys
        // below T go to the partition's own range of a spare per-point array, are sorted there by the CTA ("flip"
        // bitonic network, slots past c_less act as +inf) and summed from there. Same values in the same order.
        uint32_t *g = lpr_spill + off + lo;
        if (tid == 0)
            sm.u[5] = 0u;
        __syncthreads();
        for (uint32_t base = lo; base < hi; base += kFitThreads * kFitUnroll)
        {
#pragma unroll
            for (int h = 0; h < kFitUnroll; ++h)
            {
                const uint32_t i = base + h * kFitThreads + tid;
                const uint32_t k = i < hi ? __ldg(&zk[i]) : 0u;
                const bool take = i < hi && (!use_kmin || k > kmin) && k < kT;
                const uint32_t ballot = __ballot_sync(kFullMask, take);
                if (ballot == 0u)
                    continue;
                uint32_t wbase = 0u;
                if (lane_id() == 0)
                    wbase = atomicAdd(&sm.u[5], static_cast<uint32_t>(__popc(ballot)));
                wbase = __shfl_sync(kFullMask, wbase, 0);
                if (take)
                    g[wbase + __popc(ballot & lanemask_lt())] = k;
            }
        }
        __syncthreads();
        uint32_t n_pad2 = 2u;
        while (n_pad2 < c_less)
            n_pad2 <<= 1;
        for (uint32_t kk = 2u; kk <= n_pad2; kk <<= 1)
            for (uint32_t jj = kk >> 1; jj > 0u; jj >>= 1)
            {
                const uint32_t lj = 31u - __clz(jj);
                for (uint32_t t = tid; t < (n_pad2 >> 1); t += kFitThreads)
                {
                    uint32_t i0, i1;
                    if (jj == (kk >> 1))
                    {
                        const uint32_t blk = t >> lj, o = t & (jj - 1u);
                        i0 = blk * kk + o;
                        i1 = blk * kk + kk - 1u - o;
                    }
                    else
                    {
                        i0 = ((t & ~(jj - 1u)) << 1) | (t & (jj - 1u));
                        i1 = i0 | jj;
                    }
                    if (i1 < c_less)
                    {
                        const uint32_t a = g[i0], b = g[i1];
                        if (a > b)
                        {
                            g[i0] = b;
                            g[i1] = a;
                        }
                    }
                }
                __syncthreads();
            }
        if (tid == 0)
        {
            float zsum = 0.0f;
            uint32_t i = 0;
            for (; i + 8u <= c_less; i += 8u) // loads ahead of the dependent FADD chain
            {
                uint32_t v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    v[q] = g[i + q];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    zsum = __fadd_rn(zsum, ordered_to_float(v[q]));
            }
            for (; i < c_less; ++i)
                zsum = __fadd_rn(zsum, ordered_to_float(g[i]));
            const float zt = ordered_to_float(kT);
            for (uint32_t r = 0; r < remaining; ++r)
                zsum = __fadd_rn(zsum, zt);
            const float zmean = __fdiv_rn(zsum, static_cast<float>(n_lpr));
            sm.fl[0] = __fadd_rn(zmean, prm.initial_seed_threshold);
        }
        __syncthreads();
    }
    else
    {
    // gather the keys below T, sort them ascending (bitonic in smem)
    uint32_t n_pad = 256u; // at least one warp chunk
    while (n_pad < c_less)
        n_pad <<= 1;
    if (tid == 0)
        sm.u[5] = 0u;
    for (uint32_t i = tid; i < n_pad; i += kFitThreads)
        sm.buf[i] = 0xFFFFFFFFu;
    __syncthreads();
    for (uint32_t base = lo; base < hi; base += kFitThreads * kFitUnroll)
    {
        uint32_t k[kFitUnroll];
        bool valid[kFitUnroll];
#pragma unroll
        for (int h = 0; h < kFitUnroll; ++h)
        {
            const uint32_t i = base + h * kFitThreads + tid;
            valid[h] = i < hi;
            k[h] = valid[h] ? __ldg(&zk[i]) : 0u;
        }
#pragma unroll
        for (int h = 0; h < kFitUnroll; ++h)
        {
            const bool take = valid[h] && (!use_kmin || k[h] > kmin) && k[h] < kT;
            const uint32_t ballot = __ballot_sync(kFullMask, take);
            if (ballot == 0u)
                continue;
            uint32_t wbase = 0u;
            if (lane_id() == 0)
                wbase = atomicAdd(&sm.u[5], static_cast<uint32_t>(__popc(ballot)));
            wbase = __shfl_sync(kFullMask, wbase, 0);
            if (take)
                sm.buf[wbase + __popc(ballot & lanemask_lt())] = k[h];
        }
    }
    __syncthreads();
    // Bitonic sort in shared memory. Warp w owns elements [256 w, 256 w + 256): every compare-exchange with
    // j < 256 stays inside one warp's chunk and needs only __syncwarp; only the 15 stages with j >= 256 are
    // CTA-wide (76 of the 91 stages of an 8192-key sort run without a CTA barrier).
    {
        const uint32_t lane = lane_id();
        const uint32_t chunk = (tid >> 5) * 256u;
        auto exchange = [&](uint32_t i, uint32_t j, uint32_t kk) {
            const uint32_t ixj = i | j;
            const uint32_t a = sm.buf[i];
            const uint32_t b = sm.buf[ixj];
            const bool asc = (i & kk) == 0u;
            if ((a > b) == asc)
            {
                sm.buf[i] = b;
                sm.buf[ixj] = a;
            }
        };
        auto warp_stages = [&](uint32_t kk, uint32_t j_first) {
            if (chunk < n_pad)
                for (uint32_t j = j_first; j > 0u; j >>= 1)
                {
#pragma unroll
                    for (uint32_t r = 0; r < 4u; ++r)
                    {
                        const uint32_t t = lane + 32u * r; // pair number inside the chunk
                        exchange(chunk + (((t & ~(j - 1u)) << 1) | (t & (j - 1u))), j, kk);
                    }
                    __syncwarp();
                }
        };
        for (uint32_t kk = 2u; kk <= n_pad; kk <<= 1)
        {
            if (kk <= 256u)
            {
                warp_stages(kk, kk >> 1);
                continue;
            }
            __syncthreads();
            for (uint32_t j = kk >> 1; j >= 256u; j >>= 1)
            {
                for (uint32_t t = tid; t < (n_pad >> 1); t += kFitThreads)
                    exchange(((t & ~(j - 1u)) << 1) | (t & (j - 1u)), j, kk);
                __syncthreads();
            }
            warp_stages(kk, 128u);
        }
        __syncthreads();
    }
    // ascending sequential float sum (segmentation.cpp:189-197) — order-exact, one thread
    if (tid == 0)
    {
        float zsum = 0.0f;
        uint32_t i = 0;
        for (; i + 8u <= c_less; i += 8u) // loads ahead of the dependent FADD chain
        {
            const uint4 a = *reinterpret_cast<const uint4 *>(&sm.buf[i]);
            const uint4 b = *reinterpret_cast<const uint4 *>(&sm.buf[i + 4u]);
            zsum = __fadd_rn(zsum, ordered_to_float(a.x));
            zsum = __fadd_rn(zsum, ordered_to_float(a.y));
            zsum = __fadd_rn(zsum, ordered_to_float(a.z));
            zsum = __fadd_rn(zsum, ordered_to_float(a.w));
            zsum = __fadd_rn(zsum, ordered_to_float(b.x));
            zsum = __fadd_rn(zsum, ordered_to_float(b.y));
            zsum = __fadd_rn(zsum, ordered_to_float(b.z));
            zsum = __fadd_rn(zsum, ordered_to_float(b.w));
        }
        for (; i < c_less; ++i)
            zsum = __fadd_rn(zsum, ordered_to_float(sm.buf[i]));
        const float zt = ordered_to_float(kT);
        for (uint32_t i = 0; i < remaining; ++i)
            zsum = __fadd_rn(zsum, zt);
        const float zmean = __fdiv_rn(zsum, static_cast<float>(n_lpr));
        sm.fl[0] = __fadd_rn(zmean, prm.initial_seed_threshold);
    }
    __syncthreads();
    }
    const float zmax = sm.fl[0];
    const uint32_t kzmax = float_to_ordered(zmax);
    // seeds = sorted prefix before the first z > z_max; none found -> zero seeds (segmentation.cpp:199-216)
    const bool have_seeds = kmax > kzmax;

    // ---- fit_ground_plane iterations (segmentation.cpp:247-308) ----
    const float4 p0 = seg[lo];
    const double sx = p0.x, sy = p0.y, sz = p0.z;
    PlaneF plane;
    bool failed = !have_seeds;
    for (uint32_t it = 0; it <= prm.iterations && !failed; ++it)
    {
        // pass `it`: it == 0 selects the seeds; it >= 1 classifies with plane `it` and, unless it
        // is the last pass, accumulates the moments of the new ground set for the next fit.
        const bool last = it == prm.iterations;
        double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t cnt = 0u;
        for (uint32_t base = lo; base < hi; base += kFitThreads * kFitUnroll)
        {
            float4 pp[kFitUnroll];
#pragma unroll
            for (int h = 0; h < kFitUnroll; ++h)
            {
                const uint32_t i = base + h * kFitThreads + tid;
                if (i < hi)
                    pp[h] = __ldg(&seg[i]);
            }
#pragma unroll
            for (int h = 0; h < kFitUnroll; ++h)
            {
                const uint32_t i = base + h * kFitThreads + tid;
                if (i >= hi)
                    break;
                const float4 p = pp[h];
                bool g;
                if (it == 0u)
                {
                    const uint32_t k = float_to_ordered(p.z);
                    g = (!use_kmin || k > kmin) && k <= kzmax;
                }
                else
                    g = fit_is_ground(p, plane);
                if (last)
                    fl[i] = g ? 1u : 2u;
                else if (g)
                {
                    const double x = static_cast<double>(p.x) - sx;
                    const double y = static_cast<double>(p.y) - sy;
                    const double z = static_cast<double>(p.z) - sz;
                    acc[0] += x;
                    acc[1] += y;
                    acc[2] += z;
                    acc[3] += x * x;
                    acc[4] += x * y;
                    acc[5] += x * z;
                    acc[6] += y * y;
                    acc[7] += y * z;
                    acc[8] += z * z;
                    ++cnt;
                }
            }
        }
        if (last)
            break;
        double mom[10];
        uint32_t n_ground;
        fit_reduce_moments(sm, acc, cnt, mom, &n_ground);
        if (tid == 0)
        {
            PlaneF pl;
            const bool ok = fit_plane_from_moments(mom, n_ground, sx, sy, sz, prm.orthogonal_distance_threshold, &pl);
            sm.u[6] = ok ? 1u : 0u;
            sm.fl[1] = pl.a;
            sm.fl[2] = pl.b;
            sm.fl[3] = pl.c;
            sm.fl[4] = pl.d;
            sm.fl[5] = pl.thr;
            if (ok)
            {
                planes[it * 4 + 0] = pl.a;
                planes[it * 4 + 1] = pl.b;
                planes[it * 4 + 2] = pl.c;
                planes[it * 4 + 3] = pl.d;
            }
        }
        __syncthreads();
        failed = sm.u[6] == 0u;
        plane.a = sm.fl[1];
        plane.b = sm.fl[2];
        plane.c = sm.fl[3];
        plane.d = sm.fl[4];
        plane.thr = sm.fl[5];
        __syncthreads();
    }
    if (failed) // "Failed ground segmentation": everything becomes OBSTACLE (segmentation.cpp:251-259, 275-283)
        for (uint32_t i = lo + tid; i < hi; i += kFitThreads)
            fl[i] = 2u;
    if (tid == 0)
        *status = failed ? 2 : 0;
}

// ---------------------------------------------------------------------------------------------
// Stable compaction in tiles of 4096 points (grid = (tiles, frames)): pass 1 counts ground / obstacle
// points per tile, pass 2 adds the counts of the preceding tiles and writes. Output order = x-sorted
// order, i.e. partition 0 then 1 ..., exactly the push_back order of segmentation.cpp:331-343.
constexpr int kCompactPer = 4;
constexpr uint32_t kCompactTile = 1024u * kCompactPer;

__global__ void __launch_bounds__(1024)
seg_compact_count_kernel(const uint8_t *__restrict__ flags, BatchView bv, uint32_t max_tiles,
                         unsigned long long *__restrict__ tile_counts)
{
    __shared__ unsigned long long ws64[33];
    const uint32_t f = blockIdx.y;
    const uint32_t n = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t first = blockIdx.x * kCompactTile + threadIdx.x * kCompactPer;
    unsigned long long packed = 0ull; // ground count in the low word, obstacle count in the high word
#pragma unroll
    for (int k = 0; k < kCompactPer; ++k)
    {
        const uint8_t fl = (first + k < n) ? flags[off + first + k] : 0u;
        packed += (fl == 1u ? 1ull : 0ull) + (fl == 2u ? (1ull << 32) : 0ull);
    }
    const unsigned long long total = block_reduce_add64<1024>(packed, ws64);
    if (threadIdx.x == 0)
        tile_counts[static_cast<size_t>(f) * max_tiles + blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
seg_compact_kernel(const float4 *__restrict__ spts, const uint8_t *__restrict__ flags, BatchView bv, uint32_t max_tiles,
                   const unsigned long long *__restrict__ tile_counts, uint32_t *__restrict__ labels,
                   uint32_t *__restrict__ ground_idx, uint32_t *__restrict__ obstacle_idx,
                   float4 *__restrict__ obstacle_pts, uint32_t *__restrict__ n_ground, uint32_t *__restrict__ n_obstacle)
{
    __shared__ uint32_t ws[33];
    __shared__ unsigned long long ws64[33];
    const uint32_t f = blockIdx.y;
    const uint32_t n = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t base = blockIdx.x * kCompactTile;
    if (base >= n && !(blockIdx.x == 0u))
        return;
    const unsigned long long before = tile_prefix64<1024>(tile_counts + static_cast<size_t>(f) * max_tiles, blockIdx.x, ws64);
    const uint32_t first = base + threadIdx.x * kCompactPer;
    uint8_t fl[kCompactPer];
    uint32_t packed = 0u; // ground count in low 16 bits, obstacle count in high 16 bits (<= 4096 each)
#pragma unroll
    for (int k = 0; k < kCompactPer; ++k)
    {
        fl[k] = (first + k < n) ? flags[off + first + k] : 0u;
        packed += (fl[k] == 1u ? 1u : 0u) + (fl[k] == 2u ? 0x10000u : 0u);
    }
    uint32_t total;
    const uint32_t ex = block_exclusive_scan<1024>(packed, ws, &total);
    uint32_t g = static_cast<uint32_t>(before) + (ex & 0xFFFFu);
    uint32_t o = static_cast<uint32_t>(before >> 32) + (ex >> 16);
#pragma unroll
    for (int k = 0; k < kCompactPer; ++k)
    {
        if (fl[k] == 0u)
            continue;
        const float4 p = spts[off + first + k];
        const uint32_t orig = __float_as_uint(p.w);
        if (fl[k] == 1u)
        {
            ground_idx[off + g++] = orig;
            labels[off + orig] = kSegGround;
        }
        else
        {
            obstacle_idx[off + o] = orig;
            obstacle_pts[off + o] = p;
            ++o;
            labels[off + orig] = kSegObstacle;
        }
    }
    // the tile that holds the frame's last point (tile 0 of an empty frame) publishes the totals
    if (threadIdx.x == 0 && (base + kCompactTile >= n))
    {
        n_ground[f] = static_cast<uint32_t>(before) + (total & 0xFFFFu);
        n_obstacle[f] = static_cast<uint32_t>(before >> 32) + (total >> 16);
    }
}

} // namespace lb
