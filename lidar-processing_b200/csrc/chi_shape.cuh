// Concave outlines of the clusters of 20 points and more on the device (the Delaunay-based
// chi-shape of findOrderedConcaveOutlines, reference src/polygon_simplification.cpp:119-140 ->
// Concave-Hull/concave_hull.hpp:96-193 -> Concave-Hull/delaunator.cpp), sm_100a.
//
// The algorithm of the reference is a chain of dependent pointer updates per cluster (advancing hull, edge flips, a heap
// of boundary edges); its result depends on the insertion order and on every epsilon-guarded predicate, so the device
// runs the same chain (chi_shape.h, float64, no FMA contraction) and takes its parallelism from the batch: a 154-frame
// batch holds ~20 000 such clusters, one warp each. Inside a cluster the warp works together wherever the reference's
// result does not depend on an order: copying the points, bounding box, the three seed searches (first index among equal
// minima = lexicographic (value, index) reduction), the distances, the sort by (distance, index) (a bitonic network over
// 64-bit keys) and the scan for equally distant different points that sends a cluster to the std::sort re-enactment.
// The sweep, the flips and the erosion run on lane 0 out of the cluster's own working set, which stays in L1/L2.
//
// Scheduling: tasks are bucketed by floor(log2 n), largest bucket first, and pulled from one batch-wide counter by
// persistent warps, so the few clusters of 10 000+ points start at once and the small ones fill in behind them.
//
// Working set: 144 bytes per grouped point, cluster k of frame f at arena + 144 * (off[f] + goff[k]) (chi_layout(n).bytes
// = 141 n + 4 sqrt(n) - 138 + alignment < 144 n for n >= 20), so no prefix sum and no host round trip is needed to place
// it. The pseudo-angle hash of a cluster (ceil(sqrt(n)) words) lives in shared memory.
#pragma once

#include "chi_shape.h"
#include "common.cuh"
#include "hull.cuh"

namespace lb
{

constexpr uint32_t kChiBytesPerPoint = 144u;
constexpr uint32_t kChiSortSmem = 2048u;     // records of the CTA's staging buffer for the std::sort re-enactment (16 KB)
constexpr uint32_t kChiSmemHash = 512u;      // hash words per warp in shared memory (clusters up to 262 144 points)
constexpr int kChiWarps = 4;
constexpr uint32_t kChiBuckets = 32u;
constexpr uint32_t kChiStatTasks = 4096u;     // tasks with a diagnostics record (the largest ones come first)
constexpr uint32_t kHullErrCollinear = 8u;  // the reference throws "not triangulation" on this cluster
constexpr uint32_t kHullErrDegenerate = 16u; // every point of the cluster coincides (the reference reads out of bounds) / flip budget
constexpr uint32_t kHullErrSlot = 32u;       // closed outlines of a frame beyond its slot

struct ChiView
{
    const float4 *gpts;   // grouped points (group.cuh), frame-major
    const uint32_t *goff; // CSR offsets, frame f at [off[f] + f, off[f] + f + K]
    uint32_t *hres;       // per cluster, at its CSR position: outline vertices as cluster-local indices (open loop)
    uint32_t *hcnt;       // per cluster (same layout as goff): vertices of the CLOSED outline
    unsigned char *arena; // 144 bytes per point slot
    uint32_t *err;
};

// counts[b] = clusters of the batch with floor(log2 n) == b and n >= 20 (grid = frames)
__global__ void __launch_bounds__(256)
chi_bucket_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, const uint32_t *__restrict__ goff,
                  uint32_t *__restrict__ counts)
{
    __shared__ uint32_t s_cnt[kChiBuckets];
    const uint32_t f = blockIdx.x;
    const uint32_t K = n_clusters[f];
    const uint32_t *go = goff + bv.off[f] + f;
    if (threadIdx.x < kChiBuckets)
        s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < K; k += 256u)
    {
        const uint32_t n = go[k + 1u] - go[k];
        if (n >= kHullConcaveMin)
            atomicAdd(&s_cnt[31u - __clz(n)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kChiBuckets && s_cnt[threadIdx.x])
        atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]);
}

// task list in descending bucket order (inside a bucket in arrival order: the order of the tasks changes the schedule,
// not a result)
__global__ void __launch_bounds__(256)
chi_place_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, const uint32_t *__restrict__ goff,
                 const uint32_t *__restrict__ counts, uint32_t *__restrict__ fill, uint32_t *__restrict__ task_f,
                 uint32_t *__restrict__ task_k)
{
    __shared__ uint32_t s_base[kChiBuckets];
    const uint32_t f = blockIdx.x;
    const uint32_t K = n_clusters[f];
    const uint32_t *go = goff + bv.off[f] + f;
    if (threadIdx.x < kChiBuckets)
    {
        uint32_t before = 0u;
        for (uint32_t b = threadIdx.x + 1u; b < kChiBuckets; ++b)
            before += counts[b];
        s_base[threadIdx.x] = before;
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < K; k += 256u)
    {
        const uint32_t n = go[k + 1u] - go[k];
        if (n >= kHullConcaveMin)
        {
            const uint32_t b = 31u - __clz(n);
            const uint32_t at = s_base[b] + atomicAdd(&fill[b], 1u);
            task_f[at] = f;
            task_k[at] = k;
        }
    }
}

// lexicographic (value, index) minimum across the warp; every lane ends with the result
LB_D void chi_warp_argmin(double &v, uint32_t &i)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        const double ov = __shfl_xor_sync(kFullMask, v, d);
        const uint32_t oi = __shfl_xor_sync(kFullMask, i, d);
        if (ov < v || (ov == v && oi < i))
        {
            v = ov;
            i = oi;
        }
    }
}

LB_D double chi_warp_min(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v = fmin(v, __shfl_xor_sync(kFullMask, v, d));
    return v;
}

LB_D double chi_warp_max(double v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v = fmax(v, __shfl_xor_sync(kFullMask, v, d));
    return v;
}

// The seed triangle of chi_seed_sequential with the warp: same values, every lane holds them afterwards.
LB_D uint32_t chi_seed_warp(ChiWork &w)
{
    const uint32_t n = w.n, lane = lane_id();
    double max_x = -DBL_MAX, max_y = -DBL_MAX, min_x = DBL_MAX, min_y = DBL_MAX;
    for (uint32_t i = lane; i < n; i += 32u)
    {
        const double x = chi_px(w, i), y = chi_py(w, i);
        min_x = fmin(min_x, x); // (which of -0.0 / +0.0 survives does not reach any result: the box only enters as
        min_y = fmin(min_y, y); //  differences and sums that are squared or compared)
        max_x = fmax(max_x, x);
        max_y = fmax(max_y, y);
    }
    min_x = chi_warp_min(min_x);
    min_y = chi_warp_min(min_y);
    max_x = chi_warp_max(max_x);
    max_y = chi_warp_max(max_y);
    const double width = max_x - min_x;
    const double height = max_y - min_y;
    w.span = width * width + height * height;
    const double bx = (min_x + max_x) / 2.0, by = (min_y + max_y) / 2.0;
    double best = DBL_MAX;
    uint32_t i0 = kChiNone, i1 = kChiNone, i2 = kChiNone;
    for (uint32_t i = lane; i < n; i += 32u)
    {
        const double d = chi_dist2(chi_px(w, i), chi_py(w, i), bx, by);
        if (d < best)
        {
            i0 = i;
            best = d;
        }
    }
    chi_warp_argmin(best, i0);
    if (i0 == kChiNone)
        return kChiErrCoincident;
    const double p0x = chi_px(w, i0), p0y = chi_py(w, i0);
    best = DBL_MAX;
    for (uint32_t i = lane; i < n; i += 32u)
    {
        const double d = chi_dist2(chi_px(w, i), chi_py(w, i), p0x, p0y);
        if (i != i0 && d < best && d > 0.0)
        {
            i1 = i;
            best = d;
        }
    }
    chi_warp_argmin(best, i1);
    if (i1 == kChiNone)
        return kChiErrCoincident;
    const double p1x = chi_px(w, i1), p1y = chi_py(w, i1);
    best = DBL_MAX;
    for (uint32_t i = lane; i < n; i += 32u)
    {
        if (i == i0 || i == i1)
            continue;
        const double r = chi_circumradius2(p0x, p0y, p1x, p1y, chi_px(w, i), chi_py(w, i));
        if (r < best)
        {
            i2 = i;
            best = r;
        }
    }
    chi_warp_argmin(best, i2);
    if (!(best < DBL_MAX))
        return kChiErrCollinear;
    if (chi_ccw(p0x, p0y, p1x, p1y, chi_px(w, i2), chi_py(w, i2)))
    {
        const uint32_t t = i1;
        i1 = i2;
        i2 = t;
    }
    w.i0 = i0;
    w.i1 = i1;
    w.i2 = i2;
    w.s0x = p0x;
    w.s0y = p0y;
    w.s1x = chi_px(w, i1);
    w.s1y = chi_py(w, i1);
    w.s2x = chi_px(w, i2);
    w.s2y = chi_py(w, i2);
    chi_circumcentre(w.s0x, w.s0y, w.s1x, w.s1y, w.s2x, w.s2y, w.cx, w.cy);
    for (uint32_t i = lane; i < n; i += 32u)
    {
        const double x = chi_px(w, i), y = chi_py(w, i);
        w.dist[i] = chi_dist2(x, y, w.cx, w.cy);
        w.node[i].key = chi_hash_key(w, x, y);
    }
    __syncwarp();
    return kChiOk;
}

// Ascending sort of n (key, index) pairs in global memory by one warp ("flip" bitonic network, slots past n = +inf);
// the index breaks ties, so equal keys end in ascending index order.
LB_D void chi_warp_sort(unsigned long long *a, uint32_t *ix, uint32_t n)
{
    constexpr int kU = 4; // pairs per lane in flight: the pairs of a stage are disjoint, so their loads can all go first
    const uint32_t lane = lane_id();
    uint32_t n_pad = 2u;
    while (n_pad < n)
        n_pad <<= 1;
    for (uint32_t kk = 2u; kk <= n_pad; kk <<= 1)
        for (uint32_t jj = kk >> 1; jj > 0u; jj >>= 1)
        {
            const uint32_t lj = 31u - __clz(jj);
            const bool flip = jj == (kk >> 1);
            for (uint32_t t0 = lane; t0 < (n_pad >> 1); t0 += 32u * kU)
            {
                uint32_t p0[kU], p1[kU], xi[kU], yi[kU];
                unsigned long long x[kU], y[kU];
                bool on[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u)
                {
                    const uint32_t t = t0 + 32u * u;
                    if (flip)
                    {
                        const uint32_t blk = t >> lj, o = t & (jj - 1u);
                        p0[u] = blk * kk + o;
                        p1[u] = blk * kk + kk - 1u - o;
                    }
                    else
                    {
                        p0[u] = ((t & ~(jj - 1u)) << 1) | (t & (jj - 1u));
                        p1[u] = p0[u] | jj;
                    }
                    on[u] = t < (n_pad >> 1) && p1[u] < n;
                    if (on[u])
                    {
                        x[u] = a[p0[u]];
                        y[u] = a[p1[u]];
                        xi[u] = ix[p0[u]];
                        yi[u] = ix[p1[u]];
                    }
                }
#pragma unroll
                for (int u = 0; u < kU; ++u)
                    if (on[u] && (x[u] > y[u] || (x[u] == y[u] && xi[u] > yi[u])))
                    {
                        a[p0[u]] = y[u];
                        a[p1[u]] = x[u];
                        ix[p0[u]] = yi[u];
                        ix[p1[u]] = xi[u];
                    }
            }
            __syncwarp();
        }
}

// One step of std::__introsort_loop (chi_sort_partition_step) over a part of more than kChiSortSmem records, by one warp:
// lane 0 runs the two scans and the swaps of the unguarded Hoare partition out of two windows of the staging buffer, one
// moving up from the left end and one moving down from the right end; the warp writes a finished window back and loads
// the next one. When the windows would meet, both go back to global memory and lane 0 finishes the step there (at most
// kChiSortSmem records). Same comparisons, same swaps, same cut.
LB_D uint32_t chi_partition_step_warp(ChiKeyed *a, uint32_t first, uint32_t last, ChiKeyed *buf)
{
    constexpr uint32_t W = kChiSortSmem / 2u;
    const uint32_t lane = lane_id();
    ChiKeyed *wl = buf, *wh = buf + W;
    uint32_t kp = 0u;
    if (lane == 0u)
    {
        // __move_median_to_first(first, first + 1, mid, last - 1)
        const uint32_t ia = first + 1u, ib = first + (last - first) / 2u, ic = last - 1u;
        const uint32_t ka = a[ia].d, kb = a[ib].d, kc = a[ic].d;
        uint32_t pick;
        if (ka < kb)
            pick = (kb < kc) ? ib : ((ka < kc) ? ic : ia);
        else
            pick = (ka < kc) ? ia : ((kb < kc) ? ic : ib);
        const ChiKeyed t = a[first];
        a[first] = a[pick];
        a[pick] = t;
        kp = a[first].d;
    }
    __syncwarp();
    uint32_t lo = first + 1u, hi = last;           // lane 0's copies count
    uint32_t lo_base = first + 1u, hi_base = last - W; // last - first > 2 W, so the windows start disjoint
    bool windowed = true;
    for (uint32_t i = lane; i < W; i += 32u)
    {
        wl[i] = a[lo_base + i];
        wh[i] = a[hi_base + i];
    }
    __syncwarp();
    uint32_t phase = 0u, cut = 0u;
    while (true)
    {
        uint32_t req = 0u; // 0: done, 1: the left window is used up, 2: the right one
        if (lane == 0u)
        {
            while (true)
            {
                if (phase == 0u) // while (comp(first, pivot)) ++first;
                {
                    if (windowed && lo >= lo_base + W)
                    {
                        req = 1u;
                        break;
                    }
                    const uint32_t v = windowed ? wl[lo - lo_base].d : a[lo].d;
                    if (v < kp)
                    {
                        ++lo;
                        continue;
                    }
                    phase = 1u;
                }
                if (phase == 1u) // --last;
                {
                    --hi;
                    phase = 2u;
                }
                if (phase == 2u) // while (comp(pivot, last)) --last;
                {
                    if (windowed && hi < hi_base)
                    {
                        req = 2u;
                        break;
                    }
                    const uint32_t v = windowed ? wh[hi - hi_base].d : a[hi].d;
                    if (kp < v)
                    {
                        --hi;
                        continue;
                    }
                    phase = 3u;
                }
                if (!(lo < hi)) // return first;
                {
                    cut = lo;
                    req = 0u;
                    break;
                }
                ChiKeyed *pl = windowed ? &wl[lo - lo_base] : &a[lo];
                ChiKeyed *ph = windowed ? &wh[hi - hi_base] : &a[hi];
                const ChiKeyed t = *pl;
                *pl = *ph;
                *ph = t;
                ++lo;
                phase = 0u;
            }
        }
        __syncwarp();
        req = __shfl_sync(kFullMask, req, 0);
        if (req == 0u)
            break;
        // (lo_base, hi_base and windowed change the same way in every lane)
        if (req == 1u)
        {
            for (uint32_t i = lane; i < W; i += 32u)
                a[lo_base + i] = wl[i];
            lo_base += W;
            if (lo_base + W > hi_base)
            {
                for (uint32_t i = lane; i < W; i += 32u)
                    a[hi_base + i] = wh[i];
                windowed = false;
            }
            else
                for (uint32_t i = lane; i < W; i += 32u)
                    wl[i] = a[lo_base + i];
        }
        else
        {
            for (uint32_t i = lane; i < W; i += 32u)
                a[hi_base + i] = wh[i];
            if (hi_base < lo_base + 2u * W) // the next window down would reach into the left one
            {
                for (uint32_t i = lane; i < W; i += 32u)
                    a[lo_base + i] = wl[i];
                windowed = false;
            }
            else
            {
                hi_base -= W;
                for (uint32_t i = lane; i < W; i += 32u)
                    wh[i] = a[hi_base + i];
            }
        }
        __syncwarp();
    }
    if (windowed)
        for (uint32_t i = lane; i < W; i += 32u)
        {
            a[lo_base + i] = wl[i];
            a[hi_base + i] = wh[i];
        }
    __syncwarp();
    return __shfl_sync(kFullMask, cut, 0);
}

// std::__introsort_loop over n records in global memory by one warp: the partition steps of the large parts run through
// chi_partition_step_warp, every part of at most kChiSortSmem records is copied into the CTA's staging buffer by the
// whole warp, finished there by lane 0 (33-cycle loads instead of L2 round trips) and copied back. Same comparisons and
// moves as chi_introsort_loop. The buffer is shared by the warps of the CTA through `lock` (clusters that need it are rare).
LB_D void chi_introsort_loop_warp(ChiKeyed *a, uint32_t n, ChiKeyed *buf, int *lock)
{
    const uint32_t lane = lane_id();
    if (lane == 0u)
        while (atomicCAS(lock, 0, 1) != 0)
            __nanosleep(200);
    __syncwarp();
    uint32_t st_first[64], st_last[64], st_depth[64];
    uint32_t sp = 1u;
    st_first[0] = 0u;
    st_last[0] = n;
    st_depth[0] = chi_sort_depth_limit(n);
    while (true)
    {
        uint32_t first = 0u, last = 0u, depth = 0u;
        if (lane == 0u && sp > 0u)
        {
            --sp;
            first = st_first[sp];
            last = st_last[sp];
            depth = st_depth[sp];
        }
        first = __shfl_sync(kFullMask, first, 0);
        last = __shfl_sync(kFullMask, last, 0);
        depth = __shfl_sync(kFullMask, depth, 0);
        if (last == 0u) // (a part on the stack has last > first >= 0)
            break;
        while (last - first > 16u)
        {
            const uint32_t m = last - first;
            if (m <= kChiSortSmem)
            {
                for (uint32_t i = lane; i < m; i += 32u)
                    buf[i] = a[first + i];
                __syncwarp();
                if (lane == 0u)
                    chi_introsort_loop(buf, 0u, m, depth);
                __syncwarp();
                for (uint32_t i = lane; i < m; i += 32u)
                    a[first + i] = buf[i];
                __syncwarp();
                break;
            }
            if (depth == 0u)
            {
                if (lane == 0u)
                    chi_sort_heap_sort(a + first, m);
                __syncwarp();
                break;
            }
            --depth;
            const uint32_t cut = chi_partition_step_warp(a, first, last, buf);
            if (lane == 0u && sp < 64u)
            {
                st_first[sp] = cut;
                st_last[sp] = last;
                st_depth[sp] = depth;
                ++sp;
            }
            last = cut;
        }
    }
    __syncwarp();
    if (lane == 0u)
    {
        __threadfence_block();
        atomicExch(lock, 0);
    }
}

// One warp per cluster; persistent warps pull (frame, cluster) tasks in the order of chi_place_kernel.
__global__ void __launch_bounds__(32 * kChiWarps, 4)
chi_outline_kernel(BatchView bv, ChiView cv, const uint32_t *__restrict__ counts, const uint32_t *__restrict__ task_f,
                   const uint32_t *__restrict__ task_k, uint32_t *__restrict__ cursor, unsigned long long *__restrict__ stats)
{
    __shared__ uint32_t s_hash[kChiWarps][kChiSmemHash];
    __shared__ ChiKeyed s_sort[kChiSortSmem];
    __shared__ int s_sort_lock;
    if (threadIdx.x == 0)
        s_sort_lock = 0;
    __syncthreads();
    const uint32_t lane = lane_id();
    uint32_t T = lane < kChiBuckets ? counts[lane] : 0u;
    T = warp_reduce_add(T);
    while (true)
    {
        uint32_t g = 0u;
        if (lane == 0u)
            g = atomicAdd(cursor, 1u);
        g = __shfl_sync(kFullMask, g, 0);
        if (g >= T)
            break;
        // diagnostics (LIDAR_B200_CHI_STATS=1): per task {n, start ns, end ns, cycles of seed / sort / sweep / erosion}
        const bool rec = stats != nullptr && g < kChiStatTasks && lane == 0u;
        unsigned long long t_start = 0ull;
        long long c0k = 0, c1k = 0, c2k = 0, c3k = 0, c4k = 0;
        if (rec)
        {
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
            c0k = clock64();
        }
        const uint32_t f = task_f[g], k = task_k[g];
        const uint32_t off = bv.off[f];
        const uint32_t *go = cv.goff + off + f;
        const uint32_t c0 = go[k];
        const uint32_t n = go[k + 1u] - c0;
        unsigned char *block = cv.arena + static_cast<size_t>(kChiBytesPerPoint) * (static_cast<size_t>(off) + c0);
        const ChiLayout lay = chi_layout(n);
        ChiWork w;
        chi_bind(w, block, lay, n);
        if (w.hash_size <= kChiSmemHash)
            w.hash = s_hash[threadIdx.x >> 5];
        {
            const float4 *src = cv.gpts + off + c0;
            for (uint32_t i = lane; i < n; i += 32u)
            {
                const float4 p = __ldg(&src[i]);
                w.node[i].x = static_cast<double>(p.x);
                w.node[i].y = static_cast<double>(p.y);
                w.onb[i] = 0u;
            }
        }
        __syncwarp();
        uint32_t err = chi_seed_warp(w);
        if (rec)
            c1k = clock64();
        if (err == kChiOk)
        {
            // order of the sweep: (distance, index); the keys borrow the half-edge records, which are empty until the sweep
            // (8 n + 4 n of 96 n - 240 bytes)
            unsigned long long *skey = reinterpret_cast<unsigned long long *>(w.edge);
            uint32_t *sid = reinterpret_cast<uint32_t *>(skey + ((n + 1u) & ~1u));
            for (uint32_t i = lane; i < n; i += 32u)
            {
                skey[i] = static_cast<unsigned long long>(__double_as_longlong(w.dist[i])); // distances are >= +0.0
                sid[i] = i;
            }
            __syncwarp();
            chi_warp_sort(skey, sid, n);
            bool mixed = false;
            for (uint32_t i = lane; i + 1u < n; i += 32u)
                if (skey[i] == skey[i + 1u])
                {
                    const ChiNode &a = w.node[sid[i]], &b = w.node[sid[i + 1u]];
                    if (!(a.x == b.x && a.y == b.y) && !(chi_on_seed(w, a.x, a.y) && chi_on_seed(w, b.x, b.y)))
                        mixed = true;
                }
            mixed = __any_sync(kFullMask, mixed);
            if (!mixed)
                for (uint32_t i = lane; i < n; i += 32u)
                    w.ids[i] = sid[i];
            __syncwarp();
            if (mixed)
            {
                // two different points exactly equally far: the reference's std::sort decides their order. Records
                // {rank of the distance among the distinct distances, id} in the ORIGINAL order of the points, built from
                // the sorted keys; they borrow the half-edge records behind the keys (8 n + 4 n | 8 n | 8 n + 4 n of 96 n - 240 bytes)
                const uint32_t n2 = (n + 1u) & ~1u;
                ChiKeyed *rec_sort = reinterpret_cast<ChiKeyed *>(skey + 2u * n2);
                uint32_t carry = 0u;
                for (uint32_t base = 0; base < n; base += 32u)
                {
                    const uint32_t i = base + lane;
                    const bool fresh = i < n && i > 0u && skey[i] != skey[i - 1u];
                    const uint32_t m = __ballot_sync(kFullMask, fresh);
                    if (i < n)
                    {
                        ChiKeyed r;
                        r.d = carry + __popc(m & (0xFFFFFFFFu >> (31u - lane)));
                        r.id = sid[i];
                        rec_sort[r.id] = r;
                    }
                    carry += __popc(m);
                }
                __syncwarp();
                chi_introsort_loop_warp(rec_sort, n, s_sort, &s_sort_lock);
                // __final_insertion_sort = a stable sort by key of what the loop left behind: (rank, position) as one
                // 64-bit key through the same network as above
                unsigned long long *fkey = reinterpret_cast<unsigned long long *>(rec_sort + n2);
                uint32_t *fpos = reinterpret_cast<uint32_t *>(fkey + n2);
                for (uint32_t i = lane; i < n; i += 32u)
                {
                    fkey[i] = (static_cast<unsigned long long>(rec_sort[i].d) << 32) | i;
                    fpos[i] = i;
                }
                __syncwarp();
                chi_warp_sort(fkey, fpos, n);
                for (uint32_t i = lane; i < n; i += 32u)
                    w.ids[i] = rec_sort[fpos[i]].id;
                __syncwarp();
            }
            uint32_t h = 0u;
            if (lane == 0u)
            {
                if (rec)
                    c2k = clock64();
                err = chi_triangulate(w);
                if (rec)
                    c3k = clock64();
                if (err == kChiOk)
                    h = chi_erode_and_walk(w, cv.hres + off + c0, true);
                if (rec)
                    c4k = clock64();
            }
            err = __shfl_sync(kFullMask, err, 0);
            h = __shfl_sync(kFullMask, h, 0);
            if (lane == 0u)
                cv.hcnt[off + f + k] = err == kChiOk ? h : 0u;
        }
        else if (lane == 0u)
            cv.hcnt[off + f + k] = 0u;
        if (err != kChiOk && lane == 0u)
            atomicOr(cv.err, err == kChiErrCollinear ? kHullErrCollinear : kHullErrDegenerate);
        if (rec)
        {
            unsigned long long t_end;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
            unsigned long long *o = stats + static_cast<size_t>(g) * 8u;
            o[0] = n;
            o[1] = t_start;
            o[2] = t_end;
            o[3] = static_cast<unsigned long long>(c1k - c0k);
            o[4] = c2k ? static_cast<unsigned long long>(c2k - c1k) : 0ull;
            o[5] = c3k ? static_cast<unsigned long long>(c3k - c2k) : 0ull;
            o[6] = c4k ? static_cast<unsigned long long>(c4k - c3k) : 0ull;
            o[7] = w.n_half / 3u;
        }
        __syncwarp();
    }
}

} // namespace lb
