// Per-cluster point compaction on the device: the split of the obstacle cloud
// by cluster label that Processor::process does on the host right after Clusterer::cluster
// (reference src/processor.cpp:180-200): clustered_obstacle_cloud[label].emplace_back(x, y, z) for
// i ascending, INVALID points skipped, empty clouds erased (there are none: labels are dense).
//
// Result per frame = CSR: cluster k owns grouped points [offset[k], offset[k+1]) of the frame, in
// ascending obstacle-cloud index (the reference's push order); offset[K] = number of valid points.
// It is a stable sort of (label, index) pairs: keys from the labels, the LSD radix sort of
// radix_sort.cuh (stable), then one pass that gathers the points and marks the segment heads.
#pragma once

#include "common.cuh"

namespace lb
{

// key = label, INVALID (-1) -> K (sorts behind every cluster); val = obstacle-cloud index
__global__ void __launch_bounds__(256)
group_keys_kernel(BatchView bv, const int32_t *__restrict__ labels, const uint32_t *__restrict__ n_clusters,
                  uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t K = n_clusters[f];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    {
        const int32_t l = labels[off + i];
        keys[off + i] = l < 0 ? K : static_cast<uint32_t>(l);
        vals[off + i] = i;
    }
}

// offsets of frame f live at goff[off[f] + f .. + K + 1) (one spare slot per frame for offset[K])
__global__ void __launch_bounds__(256)
group_emit_kernel(const float4 *__restrict__ pts, BatchView bv, const uint32_t *__restrict__ skeys,
                  const uint32_t *__restrict__ svals, const uint32_t *__restrict__ n_clusters,
                  float4 *__restrict__ gpts, uint32_t *__restrict__ gidx, uint32_t *__restrict__ goff)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t K = n_clusters[f];
    uint32_t *fo = goff + off + f;
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        if (m == 0u)
            fo[0] = 0u;
        else if (skeys[off + m - 1u] != K) // no INVALID point: the valid ones end at m
            fo[K] = m;
    }
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x)
    {
        const uint32_t k = skeys[off + r];
        const uint32_t i = svals[off + r];
        if (r == 0u || skeys[off + r - 1u] != k)
            fo[k] = r; // every label 0..K-1 occurs (labels are dense), k == K is the first INVALID point
        if (k < K)
        {
            const float4 p = __ldg(&pts[off + i]);
            gpts[off + r] = make_float4(p.x, p.y, p.z, 1.0f); // pcl::PointXYZ(x, y, z): data[3] = 1.0f
            gidx[off + r] = i;
        }
    }
}

} // namespace lb
