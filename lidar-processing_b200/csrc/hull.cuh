// Ordered convex outlines per cluster on the device (device part), sm_100a.
//
// What it re-enacts, bit for bit (float32, no FMA contraction, same predicates):
//   * geom::constructAndrewMonotoneChainConvexHull (reference Convex-Hull/convex_hull.hpp:212-281):
//     std::sort by Point::operator< (y-major, then x; convex_hull.hpp:51-61), the lower / upper stack
//     scans with `getOrientation(...) != COUNTERCLOCKWISE` pops, and the index map-back "first j with
//     sorted_points[h] == points[j]" under the epsilon equality of convex_hull.hpp:63-73;
//   * geom::constructChanConvexHull (convex_hull.hpp:366-424): ceil(sqrt(n)) contiguous subsets
//     (partitionVector, :337-364), a monotone chain per subset, a Jarvis march (:283-335) over the
//     concatenated sub-hulls;
//   * the callers' policy: findOrderedConvexOutlines uses CHAN above 1000 points and the monotone chain
//     otherwise (reference src/polygon_simplification.cpp:55-64); findOrderedConcaveOutlines uses the
//     monotone chain below 20 points (:100-118) and the Delaunay-based concave hull from 20 points on —
//     in mode kHullModeConcaveSmall that part stays on the host (BASELINE north star) and such clusters are
//     reported as "host"; in mode kHullModeConcave chi_shape.cuh computes it on the device.
//
// Work decomposition: every monotone chain (a cluster up to 1000 points, or one CHAN subset) is a task.
// Sort: persistent warps pull tasks from a batch-wide counter and run a warp-wide bitonic network over
// 64-bit (y, x) keys with the original index as tie-break in shared memory. Scan: the stack scan is an
// inherently sequential chain, so it runs one THREAD per task (about 2000 tasks per frame) over the
// sorted keys; the map-back reads the head of the run of equal points. The Jarvis fold is a
// warp-parallel re-enactment that evaluates every predicate against the same operands as the sequential
// loop, on merged points staged in shared memory. Equal (x, y) pairs are value-identical, so the
// unstable std::sort needs no replay.
//
// Envelope: distinct y (and x) values of a cluster are either equal or at least FLT_EPSILON apart
// (true for |v| >= 1 and for mm-quantised LiDAR returns); otherwise the reference's operator< is not a
// strict weak order and std::sort's result is unspecified. Inside the envelope the reference's epsilon
// equality is plain equality. A chain outside the envelope raises kHullErrEnvelope (status LIDAR_B200_ERR_INPUT). CHAN subsets must fit a warp's buffers:
// clusters up to ~1.04 M points; larger ones raise the error flag.
#pragma once

#include "common.cuh"

namespace lb
{

constexpr uint32_t kHullWarpCap = 1024u;        // points one warp sorts in shared memory
constexpr uint32_t kHullMonotoneMax = 1000u;    // cluster_points.size() > 1000 -> CHAN (polygon_simplification.cpp:55)
constexpr uint32_t kHullConcaveMin = 20u;       // cluster.size() < 20 -> monotone chain (polygon_simplification.cpp:100)
constexpr int kHullWarps = 8;                   // warps per CTA of hull_chain_kernel, 12 KB of dynamic shared memory each
constexpr uint32_t kHullModeConvex = 0u;        // findOrderedConvexOutlines
constexpr uint32_t kHullModeConcaveSmall = 1u;  // the convex branch of findOrderedConcaveOutlines
constexpr uint32_t kHullModeConcave = 2u;       // findOrderedConcaveOutlines: that branch + the chi-shape from 20 points on (chi_shape.cuh)
constexpr uint32_t kHullThreadScanMax = 192u;   // tasks up to this size are scanned one thread per task, larger ones by lane 0 of the sorting warp
constexpr uint32_t kHullErrOverflow = 1u, kHullErrSubset = 2u, kHullErrJarvis = 4u;
constexpr uint32_t kHullErrEnvelope = 64u; // two coordinates of a chain differ by less than FLT_EPSILON (see the envelope above)

struct __align__(16) HullWarpSmem
{
    unsigned long long key[kHullWarpCap];  // (ordered y) << 32 | ordered x, sorted ascending; then the points' float2 bits
    uint16_t idx[kHullWarpCap];            // original index of the sorted point (ties in key: ascending)
    uint16_t stack[2u * kHullWarpCap];     // hull_indices(2 * n) of convex_hull.hpp:222 (tasks scanned by their warp)
};

struct HullView
{
    const float4 *gpts;   // grouped points (group.cuh), frame-major
    const uint32_t *goff; // CSR offsets, frame f at [off[f] + f, off[f] + f + K]
    const uint32_t *gidx; // obstacle-cloud index of every grouped point
    uint32_t *hres;       // per cluster, at its CSR position: hull vertices as cluster-local indices
    uint32_t *hcnt;       // per cluster (same layout as goff): number of hull vertices; scanned in place
    uint32_t *sub_idx;    // CHAN: sub-hull vertices (cluster-local), at the subset's own range
    uint32_t *sub_cnt;    // CHAN: vertices per subset, at cluster start + subset number
    uint32_t *mrg_idx;    // CHAN: merged_indices
    float2 *mrg_xy;       // CHAN: merged_points
    unsigned long long *skey; // per task, at its own point range: the points in sorted order (float2 bits, x low)
    uint32_t *sidx;       // original index (inside the task's range) of every sorted point
    uint32_t *stk;        // two words per point: hull_indices(2 * n) of convex_hull.hpp:222
    uint32_t *err;        // error bits
};

// crossProduct (convex_hull.hpp:76-84): x1 * y2 - x2 * y1, every operation rounded on its own.
LB_D float hull_cross(float p1x, float p1y, float p2x, float p2y, float p3x, float p3y)
{
    const float x1 = __fsub_rn(p2x, p1x);
    const float y1 = __fsub_rn(p2y, p1y);
    const float x2 = __fsub_rn(p3x, p1x);
    const float y2 = __fsub_rn(p3y, p1y);
    return __fsub_rn(__fmul_rn(x1, y2), __fmul_rn(x2, y1));
}

// sort key -> the point itself, as the bits of float2{x, y} (x in the low word); a bijection, so runs of equal
// points stay runs of equal words
LB_D unsigned long long hull_key_to_point(unsigned long long k)
{
    const uint32_t xb = __float_as_uint(ordered_to_float(static_cast<uint32_t>(k)));
    const uint32_t yb = __float_as_uint(ordered_to_float(static_cast<uint32_t>(k >> 32)));
    return (static_cast<unsigned long long>(yb) << 32) | xb;
}

LB_D float2 hull_decode(unsigned long long point_bits)
{
    return make_float2(__uint_as_float(static_cast<uint32_t>(point_bits)), __uint_as_float(static_cast<uint32_t>(point_bits >> 32)));
}

// Ascending sort of n <= kHullWarpCap (key, original index) pairs by one warp ("flip" bitonic network: slots
// past n act as +inf). The index is the tie-break, so the first element of a run of equal points is the
// one with the smallest original index.
LB_D void hull_warp_sort(unsigned long long *a, uint16_t *ix, uint32_t n)
{
    const uint32_t lane = lane_id();
    uint32_t n_pad = 2u;
    while (n_pad < n)
        n_pad <<= 1;
    for (uint32_t kk = 2u; kk <= n_pad; kk <<= 1)
        for (uint32_t jj = kk >> 1; jj > 0u; jj >>= 1)
        {
            const uint32_t lj = 31u - __clz(jj); // jj is a power of two
            for (uint32_t t = lane; t < (n_pad >> 1); t += 32u)
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
                if (i1 < n)
                {
                    const unsigned long long x = a[i0], y = a[i1];
                    const uint16_t xi = ix[i0], yi = ix[i1];
                    if (x > y || (x == y && xi > yi))
                    {
                        a[i0] = y;
                        a[i1] = x;
                        ix[i0] = yi;
                        ix[i1] = xi;
                    }
                }
            }
            __syncwarp();
        }
}

// Number of CHAN subsets of a cluster of n points: static_cast<int>(std::ceil(std::sqrt(n))) (convex_hull.hpp:376).
LB_D uint32_t hull_chan_subsets(uint32_t n)
{
    return static_cast<uint32_t>(ceil(sqrt(static_cast<double>(n))));
}

// Task list of a frame (one CTA per frame): one task per monotone chain to run — a whole cluster, or one CHAN
// subset of a cluster above 1000 points — so that the warps of hull_chain_kernel pull equally small pieces of
// work. Tasks of frame f live at task_k / task_s [off[f] ..) (a cluster owns at most as many tasks as points);
// n_tasks[f] = their number. Clusters that run nothing here (mode 1, from 20 points on) get 0 vertices.
__global__ void __launch_bounds__(256)
hull_tasks_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, HullView hv, uint32_t mode,
                  uint32_t *__restrict__ task_k, uint32_t *__restrict__ task_s, uint32_t *__restrict__ n_tasks)
{
    __shared__ uint32_t ws[9];
    __shared__ uint32_t carry;
    const uint32_t f = blockIdx.x;
    const uint32_t off = bv.off[f];
    const uint32_t K = n_clusters[f];
    const uint32_t *go = hv.goff + off + f;
    uint32_t *hc = hv.hcnt + off + f;
    if (threadIdx.x == 0)
        carry = 0u;
    __syncthreads();
    for (uint32_t k0 = 0; k0 < K; k0 += 256u)
    {
        const uint32_t k = k0 + threadIdx.x;
        uint32_t cnt = 0u;
        if (k < K)
        {
            const uint32_t n = go[k + 1u] - go[k];
            if (mode == kHullModeConvex)
            {
                cnt = n <= kHullMonotoneMax ? 1u : hull_chan_subsets(n);
                if (n > kHullMonotoneMax && n / cnt + 1u > kHullWarpCap)
                {
                    atomicOr(hv.err, kHullErrSubset);
                    cnt = 0u;
                }
            }
            else
                cnt = n < kHullConcaveMin ? 1u : 0u;
            if (cnt == 0u && !(mode == kHullModeConcave && n >= kHullConcaveMin)) // (those belong to chi_outline_kernel)
                hc[k] = 0u;
        }
        uint32_t total;
        const uint32_t base = block_exclusive_scan<256>(cnt, ws, &total) + carry;
        for (uint32_t s = 0; s < cnt; ++s)
        {
            task_k[off + base + s] = k;
            task_s[off + base + s] = s;
        }
        __syncthreads();
        if (threadIdx.x == 0)
            carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        n_tasks[f] = carry;
}

// task_base[f] = tasks of the frames before f, task_base[F] = all tasks (one CTA).
__global__ void __launch_bounds__(256)
hull_task_base_kernel(const uint32_t *__restrict__ n_tasks, uint32_t frames, uint32_t *__restrict__ task_base)
{
    __shared__ uint32_t ws[9];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0)
        carry = 0u;
    __syncthreads();
    for (uint32_t f0 = 0; f0 < frames; f0 += 256u)
    {
        const uint32_t f = f0 + threadIdx.x;
        const uint32_t v = f < frames ? n_tasks[f] : 0u;
        uint32_t total;
        const uint32_t excl = block_exclusive_scan<256>(v, ws, &total) + carry;
        if (f < frames)
            task_base[f] = excl;
        __syncthreads();
        if (threadIdx.x == 0)
            carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        task_base[frames] = carry;
}

struct HullTask
{
    uint32_t f, k, pt0, start, size; // frame, cluster, first grouped point of the cluster (absolute), range inside it
    bool subset;                     // a CHAN subset (else the whole cluster)
    uint32_t s;
};

// batch-wide task number -> task (binary search over the per-frame prefix)
LB_D HullTask hull_task(uint32_t g, const BatchView &bv, const HullView &hv, const uint32_t *__restrict__ task_base,
                        const uint32_t *__restrict__ task_k, const uint32_t *__restrict__ task_s)
{
    uint32_t lo = 0u, hi = bv.frames - 1u;
    while (lo < hi) // last frame whose base is <= g
    {
        const uint32_t mid = (lo + hi + 1u) >> 1;
        if (task_base[mid] <= g)
            lo = mid;
        else
            hi = mid - 1u;
    }
    HullTask t;
    t.f = lo;
    const uint32_t off = bv.off[lo];
    const uint32_t local = g - task_base[lo];
    t.k = task_k[off + local];
    t.s = task_s[off + local];
    const uint32_t *go = hv.goff + off + lo;
    const uint32_t c0 = go[t.k];
    const uint32_t n = go[t.k + 1u] - c0;
    t.pt0 = off + c0;
    t.subset = n > kHullMonotoneMax;
    if (!t.subset)
    {
        t.start = 0u;
        t.size = n;
    }
    else
    {
        // partitionVector (convex_hull.hpp:337-364): the first n % S subsets hold one point more
        const uint32_t S = hull_chan_subsets(n);
        const uint32_t per = n / S, rem = n % S;
        t.start = t.s * per + min(t.s, rem);
        t.size = per + (t.s < rem ? 1u : 0u);
    }
    return t;
}

// Persistent warps pull tasks from one batch-wide counter and sort the task's points by Point::operator<
// (convex_hull.hpp:51-61, 224-226): keys and original indices go to the task's own range of skey / sidx.
__global__ void __launch_bounds__(32 * kHullWarps)
hull_sort_kernel(BatchView bv, HullView hv, const uint32_t *__restrict__ task_k, const uint32_t *__restrict__ task_s,
                 const uint32_t *__restrict__ task_base, uint32_t *__restrict__ cursor)
{
    extern __shared__ __align__(16) unsigned char hull_smem[];
    HullWarpSmem &ws = reinterpret_cast<HullWarpSmem *>(hull_smem)[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const uint32_t T = task_base[bv.frames];
    while (true)
    {
        uint32_t g = 0u;
        if (lane == 0u)
            g = atomicAdd(cursor, 1u);
        g = __shfl_sync(kFullMask, g, 0);
        if (g >= T)
            break;
        const HullTask t = hull_task(g, bv, hv, task_base, task_k, task_s);
        if (t.size < 3u)
            continue; // convex_hull.hpp:217-220: no hull, nothing to sort
        const float4 *pts = hv.gpts + t.pt0 + t.start;
        for (uint32_t i = lane; i < t.size; i += 32u)
        {
            const float4 p = __ldg(&pts[i]);
            ws.key[i] = (static_cast<unsigned long long>(float_to_ordered(p.y)) << 32) | float_to_ordered(p.x);
            ws.idx[i] = static_cast<uint16_t>(i);
        }
        __syncwarp();
        hull_warp_sort(ws.key, ws.idx, t.size);
        for (uint32_t i = lane; i < t.size; i += 32u) // the scans read points, not keys
            ws.key[i] = hull_key_to_point(ws.key[i]);
        __syncwarp();
        {
            // the envelope, checked where it can be violated: neighbours of the sorted chain whose y values differ by
            // less than FLT_EPSILON without being equal (the reference's operator< then looks at x although y differs,
            // convex_hull.hpp:51-61), or whose y values are equal and x values that close (its operator== then merges two
            // different points, :63-73). The outline delivered is the one of exact comparisons; the status says that the
            // reference's own answer is unspecified here.
            bool outside = false;
            for (uint32_t i = lane; i + 1u < t.size; i += 32u)
            {
                const float2 p = hull_decode(ws.key[i]), q = hull_decode(ws.key[i + 1u]);
                const float dy = q.y - p.y, dx = q.x - p.x;
                outside |= (dy > 0.0f && dy < 1.1920929e-07f) || (dy == 0.0f && dx > 0.0f && dx < 1.1920929e-07f);
            }
            if (__any_sync(kFullMask, outside) && lane == 0u)
                atomicOr(hv.err, kHullErrEnvelope);
        }
        if (t.size <= kHullThreadScanMax)
        {
            unsigned long long *sk = hv.skey + t.pt0 + t.start;
            uint32_t *si = hv.sidx + t.pt0 + t.start;
            for (uint32_t i = lane; i < t.size; i += 32u)
            {
                sk[i] = ws.key[i];
                si[i] = ws.idx[i];
            }
            __syncwarp();
            continue;
        }
        // a long chain: scanned right here by lane 0 out of shared memory (the same loops as hull_scan_chain_kernel)
        const uint32_t n = t.size;
        uint32_t k = 0u;
        if (lane == 0u)
        {
            float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
            for (uint32_t i = 0; i < n; ++i)
            {
                const float2 p = hull_decode(ws.key[i]);
                while (k >= 2u && !(hull_cross(a.x, a.y, b.x, b.y, p.x, p.y) > 0.0f))
                {
                    --k;
                    b = a;
                    if (k >= 2u)
                        a = hull_decode(ws.key[ws.stack[k - 2u]]);
                }
                ws.stack[k++] = static_cast<uint16_t>(i);
                a = b;
                b = p;
            }
            const uint32_t floor_k = k + 1u;
            for (uint32_t i = n - 1u; i-- > 0u;)
            {
                const float2 p = hull_decode(ws.key[i]);
                while (k >= floor_k && !(hull_cross(a.x, a.y, b.x, b.y, p.x, p.y) > 0.0f))
                {
                    --k;
                    b = a;
                    if (k >= 2u)
                        a = hull_decode(ws.key[ws.stack[k - 2u]]);
                }
                ws.stack[k++] = static_cast<uint16_t>(i);
                a = b;
                b = p;
            }
        }
        __syncwarp();
        k = __shfl_sync(kFullMask, k, 0);
        uint32_t h = k - 1u;
        if (h > n)
        {
            if (lane == 0u)
                atomicOr(hv.err, kHullErrOverflow);
            h = n;
        }
        uint32_t *out = (t.subset ? hv.sub_idx : hv.hres) + t.pt0 + t.start;
        for (uint32_t v = lane; v < h; v += 32u)
        {
            uint32_t pos = ws.stack[v];
            const unsigned long long kv = ws.key[pos];
            while (pos > 0u && ws.key[pos - 1u] == kv)
                --pos;
            out[v] = t.start + ws.idx[pos];
        }
        if (lane == 0u)
        {
            if (t.subset)
                hv.sub_cnt[t.pt0 + t.s] = h;
            else
                hv.hcnt[bv.off[t.f] + t.f + t.k] = h;
        }
        __syncwarp();
    }
}

// Thread per task: the lower / upper stack scans of constructAndrewMonotoneChainConvexHull (convex_hull.hpp:
// 227-251, COUNTERCLOCKWISE, OPEN) over the sorted points, then the map-back to original indices. The scan is
// a sequential chain per task; the tasks are what runs in parallel.
__global__ void __launch_bounds__(128)
hull_scan_chain_kernel(BatchView bv, HullView hv, const uint32_t *__restrict__ task_k, const uint32_t *__restrict__ task_s,
                       const uint32_t *__restrict__ task_base)
{
    const uint32_t T = task_base[bv.frames];
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < T; g += gridDim.x * blockDim.x)
    {
    const HullTask t = hull_task(g, bv, hv, task_base, task_k, task_s);
    const uint32_t n = t.size;
    if (n > kHullThreadScanMax)
        continue; // scanned by the warp that sorted it
    uint32_t h = 0u;
    if (n >= 3u)
    {
        const unsigned long long *sk = hv.skey + t.pt0 + t.start;
        uint32_t *st = hv.stk + 2ull * (t.pt0 + t.start);
        uint32_t k = 0u;
        float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f); // sorted_points[hull[k-2]], [k-1]
        for (uint32_t i = 0; i < n; ++i) // lower chain, convex_hull.hpp:229-237
        {
            const float2 p = hull_decode(sk[i]);
            while (k >= 2u && !(hull_cross(a.x, a.y, b.x, b.y, p.x, p.y) > 0.0f))
            {
                --k;
                b = a;
                if (k >= 2u)
                    a = hull_decode(sk[st[k - 2u]]);
            }
            st[k++] = i;
            a = b;
            b = p;
        }
        const uint32_t floor_k = k + 1u;
        for (uint32_t i = n - 1u; i-- > 0u;) // upper chain, convex_hull.hpp:240-248
        {
            const float2 p = hull_decode(sk[i]);
            while (k >= floor_k && !(hull_cross(a.x, a.y, b.x, b.y, p.x, p.y) > 0.0f))
            {
                --k;
                b = a;
                if (k >= 2u)
                    a = hull_decode(sk[st[k - 2u]]);
            }
            st[k++] = i;
            a = b;
            b = p;
        }
        h = k - 1u; // hull_indices.resize(k - 1)
        if (h > n)
        {
            atomicOr(hv.err, kHullErrOverflow);
            h = n;
        }
        // map back: first j with sorted_points[hull] == points[j] (convex_hull.hpp:254-265). Under the envelope of
        // this file the epsilon equality (convex_hull.hpp:63-73) is plain equality of both coordinates, the points
        // equal to a hull vertex are one run of the sorted order, and the run starts with the smallest index.
        const uint32_t *si = hv.sidx + t.pt0 + t.start;
        uint32_t *out = (t.subset ? hv.sub_idx : hv.hres) + t.pt0 + t.start;
        for (uint32_t v = 0; v < h; ++v)
        {
            uint32_t pos = st[v];
            const unsigned long long kv = sk[pos];
            while (pos > 0u && sk[pos - 1u] == kv)
                --pos;
            out[v] = t.start + si[pos];
        }
    }
    if (t.subset)
        hv.sub_cnt[t.pt0 + t.s] = h;
    else
        hv.hcnt[bv.off[t.f] + t.f + t.k] = h;
    }
}

// CTA per cluster above 1000 points: the rest of constructChanConvexHull after the per-subset chains. The
// merged points are staged in shared memory (kHullMergeCap of them; more spill to the global copy).
constexpr uint32_t kHullMergeCap = 3584u;
__global__ void __launch_bounds__(256)
hull_chan_merge_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, HullView hv)
{
    __shared__ float2 s_xy[kHullMergeCap];
    __shared__ uint32_t s_idx[kHullMergeCap];
    __shared__ uint32_t scan_ws[9];
    __shared__ uint32_t s_carry;
    const uint32_t f = blockIdx.y;
    const uint32_t off = bv.off[f];
    const uint32_t K = n_clusters[f];
    const uint32_t *go = hv.goff + off + f;
    uint32_t *hc = hv.hcnt + off + f;
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    for (uint32_t k = blockIdx.x; k < K; k += gridDim.x)
    {
        const uint32_t c0 = go[k];
        const uint32_t n = go[k + 1u] - c0;
        if (n <= kHullMonotoneMax)
            continue; // uniform across the CTA
        const uint32_t S = hull_chan_subsets(n);
        const uint32_t per = n / S, rem = n % S;
        if (per + 1u > kHullWarpCap)
            continue; // flagged by hull_tasks_kernel
        const float4 *pts = hv.gpts + off + c0;
        const uint32_t *sub_idx = hv.sub_idx + off + c0;
        const uint32_t *sub_cnt = hv.sub_cnt + off + c0;
        uint32_t *mrg_idx = hv.mrg_idx + off + c0;
        float2 *mrg_xy = hv.mrg_xy + off + c0;
        // merged_points / merged_indices: the sub-hulls end to end in subset order (convex_hull.hpp:392-406)
        __syncthreads();
        if (tid == 0)
            s_carry = 0u;
        __syncthreads();
        for (uint32_t s0 = 0; s0 < S; s0 += 256u)
        {
            const uint32_t s = s0 + tid;
            const uint32_t cnt = s < S ? sub_cnt[s] : 0u;
            uint32_t tile_total;
            const uint32_t base = block_exclusive_scan<256>(cnt, scan_ws, &tile_total) + s_carry;
            if (s < S)
            {
                const uint32_t start = s * per + min(s, rem);
                for (uint32_t v = 0; v < cnt; ++v)
                {
                    const uint32_t li = sub_idx[start + v];
                    const float4 p = __ldg(&pts[li]);
                    if (base + v < kHullMergeCap)
                    {
                        s_idx[base + v] = li;
                        s_xy[base + v] = make_float2(p.x, p.y);
                    }
                    else
                    {
                        mrg_idx[base + v] = li;
                        mrg_xy[base + v] = make_float2(p.x, p.y);
                    }
                }
            }
            __syncthreads();
            if (tid == 0)
                s_carry += tile_total;
            __syncthreads();
        }
        const uint32_t m = s_carry;
        if (tid >= 32u)
            continue; // warp 0 marches; the others wait at the barrier of the next cluster
        auto mxy = [&](uint32_t i) -> float2 { return i < kHullMergeCap ? s_xy[i] : mrg_xy[i]; };
        auto midx = [&](uint32_t i) -> uint32_t { return i < kHullMergeCap ? s_idx[i] : mrg_idx[i]; };
        // constructJarvisMarchConvexHull(merged_points) (convex_hull.hpp:283-335). The inner fold
        // `if orientation(p, i, q) == CCW then q = i` is re-enacted 32 candidates at a time: the lanes test
        // against the current q, the first hit becomes q and only the lanes behind it are tested again.
        uint32_t h = 0u;
        if (m >= 3u)
        {
            float bx = 0.f; // first index holding the minimum x (strict '<' scan)
            uint32_t bi = 0xFFFFFFFFu;
            for (uint32_t i = lane; i < m; i += 32u)
            {
                const float x = mxy(i).x;
                if (bi == 0xFFFFFFFFu || x < bx)
                {
                    bx = x;
                    bi = i;
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
            {
                const float ox = __shfl_xor_sync(kFullMask, bx, d);
                const uint32_t oi = __shfl_xor_sync(kFullMask, bi, d);
                if (oi != 0xFFFFFFFFu && (bi == 0xFFFFFFFFu || ox < bx || (ox == bx && oi < bi)))
                {
                    bx = ox;
                    bi = oi;
                }
            }
            const uint32_t leftmost = bi;
            uint32_t *out = hv.hres + off + c0;
            uint32_t p = leftmost;
            bool bad = false;
            do
            {
                if (h >= n || h > m)
                {
                    bad = true; // the reference would not terminate either
                    break;
                }
                if (lane == 0u)
                    out[h] = midx(p);
                ++h;
                const float2 pp = mxy(p);
                uint32_t q = p + 1u == m ? 0u : p + 1u;
                float2 pq = mxy(q);
                for (uint32_t i0 = 0; i0 < m; i0 += 32u)
                {
                    const uint32_t i = i0 + lane;
                    const float2 pi = i < m ? mxy(i) : pp;
                    uint32_t todo = kFullMask;
                    while (true)
                    {
                        const bool ccw = i < m && hull_cross(pp.x, pp.y, pi.x, pi.y, pq.x, pq.y) > 0.0f;
                        const uint32_t bm = __ballot_sync(kFullMask, ccw) & todo;
                        if (bm == 0u)
                            break;
                        const int first = __ffs(bm) - 1;
                        q = i0 + static_cast<uint32_t>(first);
                        pq.x = __shfl_sync(kFullMask, pi.x, first);
                        pq.y = __shfl_sync(kFullMask, pi.y, first);
                        todo = first == 31 ? 0u : (kFullMask << (first + 1));
                        if (todo == 0u)
                            break;
                    }
                }
                p = q;
            } while (p != leftmost);
            if (bad)
            {
                if (lane == 0u)
                    atomicOr(hv.err, kHullErrJarvis);
                h = 0u;
            }
        }
        if (lane == 0u)
            hc[k] = h;
    }
}

// hcnt -> exclusive offsets in place, hoff[K] = vertices of the frame; hne[k] = non-empty outlines before
// cluster k, hne[K] = their number (one CTA per frame).
__global__ void __launch_bounds__(256)
hull_scan_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, uint32_t *__restrict__ hcnt,
                 uint32_t *__restrict__ hne_all, uint32_t *__restrict__ n_vertices, uint32_t *__restrict__ n_outlines,
                 const uint32_t *__restrict__ slot_points, uint32_t *__restrict__ err)
{
    __shared__ uint32_t ws[9];
    __shared__ uint32_t carry, carry_ne;
    const uint32_t f = blockIdx.x;
    const uint32_t K = n_clusters[f];
    uint32_t *hc = hcnt + bv.off[f] + f;
    uint32_t *ne = hne_all + bv.off[f] + f;
    if (threadIdx.x == 0)
    {
        carry = 0u;
        carry_ne = 0u;
    }
    __syncthreads();
    for (uint32_t k0 = 0; k0 < K; k0 += 256u)
    {
        const uint32_t k = k0 + threadIdx.x;
        const uint32_t v = k < K ? hc[k] : 0u;
        uint32_t total, total_ne;
        const uint32_t excl = block_exclusive_scan<256>(v, ws, &total) + carry;
        const uint32_t excl_ne = block_exclusive_scan<256>(v ? 1u : 0u, ws, &total_ne) + carry_ne;
        if (k < K)
        {
            hc[k] = excl;
            ne[k] = excl_ne;
        }
        __syncthreads();
        if (threadIdx.x == 0)
        {
            carry += total;
            carry_ne += total_ne;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        hc[K] = carry;
        ne[K] = carry_ne;
        n_vertices[f] = carry;
        n_outlines[f] = carry_ne;
        // A closed outline repeats its first vertex, so the outlines of a frame can outgrow the frame's slot (one
        // vertex per staged point) when nearly every point is an outline vertex: the outlines that do not fit any more
        // get 0 vertices and the status says so.
        const uint32_t cap = (slot_points[f] + 31u) & ~31u;
        if (carry > cap)
        {
            atomicOr(err, 32u /* kHullErrSlot */);
            uint32_t old_k = hc[0], acc = 0u, acc_ne = 0u;
            for (uint32_t k = 0; k < K; ++k)
            {
                const uint32_t old_next = hc[k + 1u];
                uint32_t v = old_next - old_k;
                if (acc + v > cap)
                    v = 0u;
                hc[k] = acc;
                ne[k] = acc_ne;
                acc += v;
                acc_ne += v ? 1u : 0u;
                old_k = old_next;
            }
            hc[K] = acc;
            ne[K] = acc_ne;
            n_vertices[f] = acc;
            n_outlines[f] = acc_ne;
        }
    }
}

// Outline vertices of the frame end to end: (x, y) = geom::Point<float> records plus the obstacle-cloud index.
__global__ void __launch_bounds__(256)
hull_emit_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, HullView hv, float2 *__restrict__ hxy,
                 uint32_t *__restrict__ hsrc, uint32_t closed_from)
{
    const uint32_t f = blockIdx.y;
    const uint32_t off = bv.off[f];
    const uint32_t K = n_clusters[f];
    const uint32_t *go = hv.goff + off + f;
    const uint32_t *ho = hv.hcnt + off + f;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t k = warp; k < K; k += n_warps)
    {
        const uint32_t c0 = go[k];
        const uint32_t h0 = ho[k];
        const uint32_t h = ho[k + 1u] - h0;
        // clusters of closed_from points and more carry a closed outline (getHullIndices repeats hull_start,
        // delaunator.cpp:693-707): hres holds the open loop
        const bool closed = go[k + 1u] - c0 >= closed_from;
        for (uint32_t v = lane_id(); v < h; v += 32u)
        {
            const uint32_t li = hv.hres[off + c0 + ((closed && v + 1u == h) ? 0u : v)];
            const float4 p = __ldg(&hv.gpts[off + c0 + li]);
            hxy[off + h0 + v] = make_float2(p.x, p.y);
            hsrc[off + h0 + v] = hv.gidx[off + c0 + li];
        }
    }
}

} // namespace lb
