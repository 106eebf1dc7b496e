// Output packing on the device, sm_100a: the two conversions the reference's caller
// runs on the clusters and outlines before publishing them (reference src/processor.cpp:249-272).
//
//  * colorize_kernel — convertClusteredCloudToColorizedCloud (reference src/conversions.cpp:32-60): one
//    pcl::PointXYZRGB(x, y, z, r, g, b) record per clustered point, cluster after cluster in push order.
//    The record is PCL's 32-byte layout, which convertPCLToPointCloud2 memcpy's into the message
//    (conversions.cpp:139-162, point_step = sizeof(pcl::PointXYZRGB)): floats x, y, z, 1.0f, then the
//    packed colour word b | g << 8 | r << 16 | 255 << 24 at byte 16, then 12 bytes of padding (zero here).
//    The colours themselves are the caller's: the reference draws them with std::rand() (conversions.cpp:49-51),
//    a process-wide sequence that only the host can continue, so the table comes in as an argument.
//  * marker_points_kernel — the points of convertPointXYZTypeToMarkerArray (reference src/conversions.hpp:72-120):
//    per non-empty outline a LINE_STRIP of geometry_msgs::Point {double x, y, z = 0} with the first vertex
//    appended again to close the loop (:108-117); empty outlines produce no marker (:81-84).
#pragma once

#include "common.cuh"

namespace lb
{

// warp per cluster; rgb_all holds the colour words of all frames end to end, rgb_off[f] = first cluster of frame f
__global__ void __launch_bounds__(256)
colorize_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, const float4 *__restrict__ gpts,
                const uint32_t *__restrict__ goff, const uint32_t *__restrict__ rgb_all,
                const uint32_t *__restrict__ rgb_off, uint4 *__restrict__ out)
{
    const uint32_t f = blockIdx.y;
    const uint32_t off = bv.off[f];
    const uint32_t K = n_clusters[f];
    const uint32_t *go = goff + off + f;
    const uint32_t *rgb = rgb_all + rgb_off[f];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t k = warp; k < K; k += n_warps)
    {
        const uint32_t c0 = go[k], c1 = go[k + 1u];
        const uint32_t word = (rgb[k] & 0x00FFFFFFu) | 0xFF000000u; // a = 255 (pcl::PointXYZRGB constructor)
        for (uint32_t i = c0 + lane_id(); i < c1; i += 32u)
        {
            const float4 p = __ldg(&gpts[off + i]);
            out[2ull * (off + i)] = make_uint4(__float_as_uint(p.x), __float_as_uint(p.y), __float_as_uint(p.z),
                                               __float_as_uint(1.0f));
            out[2ull * (off + i) + 1ull] = make_uint4(word, 0u, 0u, 0u);
        }
    }
}

// hoff = outline CSR of the frame, hne[k] = non-empty outlines before cluster k (hne[K] = all of them): the marker
// of outline k owns points [hoff[k] + hne[k], hoff[k+1] + hne[k+1]) of the frame's list, which starts at 2 * off[f].
__global__ void __launch_bounds__(256)
marker_points_kernel(BatchView bv, const uint32_t *__restrict__ n_clusters, const uint32_t *__restrict__ hoff_all,
                     const uint32_t *__restrict__ hne_all, const float2 *__restrict__ hxy, double *__restrict__ out)
{
    const uint32_t f = blockIdx.y;
    const uint32_t off = bv.off[f];
    const uint32_t K = n_clusters[f];
    const uint32_t *ho = hoff_all + off + f;
    const uint32_t *ne = hne_all + off + f;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t k = warp; k < K; k += n_warps)
    {
        const uint32_t h0 = ho[k];
        const uint32_t h = ho[k + 1u] - h0;
        if (h == 0u)
            continue;
        double *dst = out + 3ull * (2ull * off + h0 + ne[k]);
        for (uint32_t v = lane_id(); v <= h; v += 32u)
        {
            const float2 p = hxy[off + h0 + (v == h ? 0u : v)]; // marker.points.push_back(marker.points[0])
            dst[3u * v] = static_cast<double>(p.x);
            dst[3u * v + 1u] = static_cast<double>(p.y);
            dst[3u * v + 2u] = 0.0;
        }
    }
}

// Result emission straight into page-locked HOST memory (fetch mode 4): one launch writes, per frame, exactly the used part
// of the four result arrays - N segmentation labels, n_ground + n_obstacle indices, n_obstacle cluster labels - to the
// caller's arrays at the frame's slot offset, with coalesced 16-byte stores over PCIe. No padding crosses the bus (the
// slot-size copies moved 39 % padding), the sizes never travel to the host first (no host round trip between the last
// kernel and the transfer), and the copy engines stay free for the next chunk's upload. dst pointers are device-visible
// addresses of mapped pinned memory; a null pointer skips that array.
__global__ void __launch_bounds__(256)
emit_results_kernel(BatchView bv, const uint32_t *__restrict__ n_ground, const uint32_t *__restrict__ n_obstacle,
                    const uint32_t *__restrict__ labels, const uint32_t *__restrict__ gidx, const uint32_t *__restrict__ oidx,
                    const int32_t *__restrict__ clabels, uint32_t *dst_labels, uint32_t *dst_gidx, uint32_t *dst_oidx,
                    int32_t *dst_clabels)
{
    const uint32_t f = blockIdx.y;
    const uint32_t off = bv.off[f]; // multiple of 32 elements: 128-byte aligned on both sides when the bases are
    const uint32_t n = bv.cnt[f];
    const uint32_t cnt[4] = {n, min(n_ground[f], n), min(n_obstacle[f], n), min(n_obstacle[f], n)};
    const uint32_t *src[4] = {labels, gidx, oidx, reinterpret_cast<const uint32_t *>(clabels)};
    uint32_t *dst[4] = {dst_labels, dst_gidx, dst_oidx, reinterpret_cast<uint32_t *>(dst_clabels)};
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        if (!dst[k])
            continue;
        const uint32_t *sp = src[k] + off;
        uint32_t *dp = dst[k] + off;
        const uint32_t c = cnt[k];
        if ((reinterpret_cast<uintptr_t>(dp) & 15u) == 0u)
        {
            const uint32_t quads = c >> 2;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(sp);
            uint4 *d4 = reinterpret_cast<uint4 *>(dp);
            for (uint32_t i = t; i < quads; i += nt)
                d4[i] = s4[i];
            for (uint32_t i = (quads << 2) + t; i < c; i += nt)
                dp[i] = sp[i];
        }
        else
            for (uint32_t i = t; i < c; i += nt)
                dp[i] = sp[i];
    }
}

} // namespace lb
