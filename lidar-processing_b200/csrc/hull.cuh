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
//     that part stays on the host (BASELINE north star) and such clusters are reported as "host".
//
// Work decomposition: the sort is a warp-wide bitonic network over 64-bit (y, x) keys in shared memory
// (a subset or a cluster of the warp path has at most 1024 points); the stack scan is an inherently
// sequential chain and runs on lane 0 out of shared memory; the map-back and the Jarvis fold are
// warp-parallel re-enactments that evaluate every predicate against the same operands as the
// sequential loops. Equal (x, y) pairs are value-identical, so the unstable std::sort needs no replay.
//
// Envelope: distinct y (and x) values of a cluster are either equal or at least FLT_EPSILON apart
// (true for |v| >= 1 and for mm-quantised LiDAR returns); otherwise the reference's operator< is not a
// strict weak order and std::sort's result is unspecified. CHAN subsets must fit a warp's buffers:
// clusters up to ~1.04 M points; larger ones raise the error flag.
#pragma once

#include "common.cuh"

namespace lb
{

constexpr uint32_t kHullWarpCap = 1024u;        // points one warp sorts in shared memory
constexpr uint32_t kHullMonotoneMax = 1000u;    // cluster_points.size() > 1000 -> CHAN (polygon_simplification.cpp:55)
constexpr uint32_t kHullConcaveMin = 20u;       // cluster.size() < 20 -> monotone chain (polygon_simplification.cpp:100)
constexpr int kHullWarps = 8;                  // warps per CTA, 12 KB of dynamic shared memory each
constexpr uint32_t kHullModeConvex = 0u;        // findOrderedConvexOutlines
constexpr uint32_t kHullModeConcaveSmall = 1u;  // the convex branch of findOrderedConcaveOutlines
constexpr uint32_t kHullErrOverflow = 1u, kHullErrSubset = 2u, kHullErrJarvis = 4u;

struct __align__(16) HullWarpSmem
{
    unsigned long long key[kHullWarpCap];  // (ordered y) << 32 | ordered x, sorted ascending
    uint16_t stack[2u * kHullWarpCap];     // hull_indices(2 * n) of convex_hull.hpp:222
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

LB_D float2 hull_decode(unsigned long long k)
{
    return make_float2(ordered_to_float(static_cast<uint32_t>(k)), ordered_to_float(static_cast<uint32_t>(k >> 32)));
}

// Ascending sort of n <= kHullWarpCap keys by one warp ("flip" bitonic network: slots past n act as +inf).
LB_D void hull_warp_sort(unsigned long long *a, uint32_t n)
{
    const uint32_t lane = lane_id();
    uint32_t n_pad = 2u;
    while (n_pad < n)
        n_pad <<= 1;
    for (uint32_t kk = 2u; kk <= n_pad; kk <<= 1)
        for (uint32_t jj = kk >> 1; jj > 0u; jj >>= 1)
        {
            for (uint32_t t = lane; t < (n_pad >> 1); t += 32u)
            {
                uint32_t i0, i1;
                if (jj == (kk >> 1))
                {
                    const uint32_t blk = t / jj, o = t - blk * jj;
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
                    if (x > y)
                    {
                        a[i0] = y;
                        a[i1] = x;
                    }
                }
            }
            __syncwarp();
        }
}

// constructAndrewMonotoneChainConvexHull(points[0..n), COUNTERCLOCKWISE, OPEN) by one warp, n <= kHullWarpCap.
// Writes the hull as indices into `pts` (plus `base`) to out[0..h) and returns h (0 when n < 3); at most
// `out_cap` vertices are kept (more is reported through *err).
LB_D uint32_t hull_warp_monotone_chain(const float4 *__restrict__ pts, uint32_t n, uint32_t base, HullWarpSmem &ws,
                                       uint32_t *__restrict__ out, uint32_t out_cap, uint32_t *__restrict__ err)
{
    if (n < 3u)
        return 0u; // convex_hull.hpp:217-220
    const uint32_t lane = lane_id();
    for (uint32_t i = lane; i < n; i += 32u)
    {
        const float4 p = __ldg(&pts[i]);
        ws.key[i] = (static_cast<unsigned long long>(float_to_ordered(p.y)) << 32) | float_to_ordered(p.x);
    }
    __syncwarp();
    hull_warp_sort(ws.key, n);
    uint32_t k = 0u;
    if (lane == 0u)
    {
        float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f); // sorted_points[hull[k-2]], [k-1]
        for (uint32_t i = 0; i < n; ++i) // lower chain, convex_hull.hpp:229-237
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
        const uint32_t t = k + 1u;
        for (uint32_t i = n - 1u; i-- > 0u;) // upper chain, convex_hull.hpp:240-248
        {
            const float2 p = hull_decode(ws.key[i]);
            while (k >= t && !(hull_cross(a.x, a.y, b.x, b.y, p.x, p.y) > 0.0f))
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
    uint32_t h = k - 1u; // hull_indices.resize(k - 1)
    if (h > out_cap)
    {
        if (lane == 0u)
            atomicOr(err, kHullErrOverflow);
        h = out_cap;
    }
    // map back: first j with sorted_points[hull] == points[j] (epsilon equality, convex_hull.hpp:63-73, 254-265)
    const float eps = 1.1920928955078125e-07f;
    for (uint32_t v = 0; v < h; ++v)
    {
        const float2 s = hull_decode(ws.key[ws.stack[v]]);
        uint32_t found = 0u; // the reference leaves the sorted position when nothing matches (cannot happen)
        for (uint32_t j0 = 0; j0 < n; j0 += 32u)
        {
            const uint32_t j = j0 + lane;
            bool eq = false;
            if (j < n)
            {
                const float4 p = __ldg(&pts[j]);
                eq = fabsf(__fsub_rn(s.x, p.x)) < eps && fabsf(__fsub_rn(s.y, p.y)) < eps;
            }
            const uint32_t bm = __ballot_sync(kFullMask, eq);
            if (bm)
            {
                found = j0 + static_cast<uint32_t>(__ffs(bm) - 1);
                break;
            }
        }
        if (lane == 0u)
            out[v] = base + found;
    }
    __syncwarp();
    return h;
}

// Warp per cluster: monotone chain for the clusters the mode assigns to it. Every cluster of the frame gets
// its hcnt written here (0 for the ones left to hull_chan_kernel / the host) except the CHAN ones in convex mode.
__global__ void __launch_bounds__(32 * kHullWarps)
hull_warp_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, HullView hv, uint32_t mode)
{
    extern __shared__ __align__(16) unsigned char hull_smem[];
    HullWarpSmem *sm = reinterpret_cast<HullWarpSmem *>(hull_smem);
    const uint32_t f = blockIdx.y;
    const uint32_t off = bv.off[f];
    const uint32_t K = n_clusters[f];
    const uint32_t *go = hv.goff + off + f;
    uint32_t *hc = hv.hcnt + off + f;
    const uint32_t warp = threadIdx.x >> 5;
    HullWarpSmem &ws = sm[warp];
    for (uint32_t k = blockIdx.x * kHullWarps + warp; k < K; k += gridDim.x * kHullWarps)
    {
        const uint32_t c0 = go[k];
        const uint32_t n = go[k + 1u] - c0;
        bool run;
        if (mode == kHullModeConvex)
        {
            if (n > kHullMonotoneMax)
                continue; // hull_chan_kernel writes this cluster's count
            run = true;
        }
        else
            run = n < kHullConcaveMin; // from 20 points on: the host's concave hull, reported with 0 vertices
        uint32_t h = 0u;
        if (run)
            h = hull_warp_monotone_chain(hv.gpts + off + c0, n, 0u, ws, hv.hres + off + c0, n, hv.err);
        if (lane_id() == 0u)
            hc[k] = h;
        __syncwarp();
    }
}

// CTA per cluster above 1000 points (convex mode): constructChanConvexHull.
__global__ void __launch_bounds__(32 * kHullWarps)
hull_chan_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, HullView hv)
{
    extern __shared__ __align__(16) unsigned char hull_smem[];
    HullWarpSmem *sm = reinterpret_cast<HullWarpSmem *>(hull_smem);
    __shared__ uint32_t scan_ws[kHullWarps + 1];
    __shared__ uint32_t s_carry;
    constexpr int NT = 32 * kHullWarps;
    const uint32_t f = blockIdx.y;
    const uint32_t off = bv.off[f];
    const uint32_t K = n_clusters[f];
    const uint32_t *go = hv.goff + off + f;
    uint32_t *hc = hv.hcnt + off + f;
    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31u;
    for (uint32_t k = blockIdx.x; k < K; k += gridDim.x)
    {
        const uint32_t c0 = go[k];
        const uint32_t n = go[k + 1u] - c0;
        if (n <= kHullMonotoneMax)
            continue; // uniform across the CTA
        const float4 *pts = hv.gpts + off + c0;
        // partitionVector (convex_hull.hpp:337-364) with ceil(sqrt(n)) subsets (:376)
        const uint32_t S = static_cast<uint32_t>(ceil(sqrt(static_cast<double>(n))));
        const uint32_t per = n / S, rem = n % S;
        if (per + 1u > kHullWarpCap)
        {
            if (tid == 0)
            {
                atomicOr(hv.err, kHullErrSubset);
                hc[k] = 0u;
            }
            continue;
        }
        uint32_t *sub_idx = hv.sub_idx + off + c0;
        uint32_t *sub_cnt = hv.sub_cnt + off + c0;
        uint32_t *mrg_idx = hv.mrg_idx + off + c0;
        float2 *mrg_xy = hv.mrg_xy + off + c0;
        for (uint32_t s = warp; s < S; s += kHullWarps)
        {
            const uint32_t start = s * per + min(s, rem);
            const uint32_t size = per + (s < rem ? 1u : 0u);
            const uint32_t h = hull_warp_monotone_chain(pts + start, size, start, sm[warp], sub_idx + start, size, hv.err);
            if (lane == 0u)
                sub_cnt[s] = h;
        }
        __syncthreads();
        // merged_points / merged_indices: the sub-hulls end to end in subset order (convex_hull.hpp:392-406)
        if (tid == 0)
            s_carry = 0u;
        __syncthreads();
        for (uint32_t s0 = 0; s0 < S; s0 += NT)
        {
            const uint32_t s = s0 + tid;
            const uint32_t cnt = s < S ? sub_cnt[s] : 0u;
            uint32_t tile_total;
            const uint32_t excl = block_exclusive_scan<NT>(cnt, scan_ws, &tile_total) + s_carry;
            if (s < S)
            {
                const uint32_t start = s * per + min(s, rem);
                for (uint32_t v = 0; v < cnt; ++v)
                {
                    const uint32_t li = sub_idx[start + v];
                    const float4 p = __ldg(&pts[li]);
                    mrg_idx[excl + v] = li;
                    mrg_xy[excl + v] = make_float2(p.x, p.y);
                }
            }
            __syncthreads();
            if (tid == 0)
                s_carry += tile_total;
            __syncthreads();
        }
        const uint32_t m = s_carry;
        // constructJarvisMarchConvexHull(merged_points) by warp 0 (convex_hull.hpp:283-335). The inner fold
        // `if orientation(p, i, q) == CCW then q = i` is re-enacted 32 candidates at a time: the lanes test
        // against the current q, the first hit becomes q and only the lanes behind it are tested again.
        if (warp == 0u)
        {
            uint32_t h = 0u;
            if (m >= 3u)
            {
                uint32_t leftmost = 0u;
                {
                    // first index holding the minimum x (strict '<' scan)
                    float bx = 0.f;
                    uint32_t bi = 0xFFFFFFFFu;
                    for (uint32_t i = lane; i < m; i += 32u)
                    {
                        const float x = mrg_xy[i].x;
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
                    leftmost = bi;
                }
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
                        out[h] = mrg_idx[p];
                    ++h;
                    const float2 pp = mrg_xy[p];
                    uint32_t q = p + 1u == m ? 0u : p + 1u;
                    float2 pq = mrg_xy[q];
                    for (uint32_t i0 = 0; i0 < m; i0 += 32u)
                    {
                        const uint32_t i = i0 + lane;
                        const float2 pi = i < m ? mrg_xy[i] : pp;
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
        __syncthreads();
    }
}

// hcnt -> exclusive offsets in place, hoff[K] = vertices of the frame (one CTA per frame).
__global__ void __launch_bounds__(256)
hull_scan_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, uint32_t *__restrict__ hcnt,
                 uint32_t *__restrict__ n_vertices)
{
    __shared__ uint32_t ws[9];
    __shared__ uint32_t carry;
    const uint32_t f = blockIdx.x;
    const uint32_t K = n_clusters[f];
    uint32_t *hc = hcnt + bv.off[f] + f;
    if (threadIdx.x == 0)
        carry = 0u;
    __syncthreads();
    for (uint32_t k0 = 0; k0 < K; k0 += 256u)
    {
        const uint32_t k = k0 + threadIdx.x;
        const uint32_t v = k < K ? hc[k] : 0u;
        uint32_t total;
        const uint32_t excl = block_exclusive_scan<256>(v, ws, &total) + carry;
        if (k < K)
            hc[k] = excl;
        __syncthreads();
        if (threadIdx.x == 0)
            carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        hc[K] = carry;
        n_vertices[f] = carry;
    }
}

// Outline vertices of the frame end to end: (x, y) = geom::Point<float> records plus the obstacle-cloud index.
__global__ void __launch_bounds__(256)
hull_emit_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, HullView hv, float2 *__restrict__ hxy,
                 uint32_t *__restrict__ hsrc)
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
        for (uint32_t v = lane_id(); v < h; v += 32u)
        {
            const uint32_t li = hv.hres[off + c0 + v];
            const float4 p = __ldg(&hv.gpts[off + c0 + li]);
            hxy[off + h0 + v] = make_float2(p.x, p.y);
            hsrc[off + h0 + v] = hv.gidx[off + c0 + li];
        }
    }
}

} // namespace lb
