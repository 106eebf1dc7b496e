// Sequential selection primitives for the k-d tree order emulation, host- and device-compilable.
//
// The reference builds its k-d tree with std::nth_element over whole node structs, comparing one
// coordinate only (reference src/kdtree.hpp:195-206). LiDAR data is mm-quantised and full of
// coordinate ties, so WHICH permutation nth_element leaves behind decides the tree, the pre-order
// in which KDTree::radius_search reports neighbours (kdtree.hpp:292-341) and therefore the FIFO
// order of the clustering BFS (src/clustering.cpp:77-111). To reproduce the reference's cluster
// partition bit for bit the device re-enacts libstdc++'s algorithm step by step:
//   introselect loop, depth limit 2*floor(log2 n); median-of-three of (first+1, mid, last-1) moved
//   to first; Hoare "unguarded" partition; keep the side holding nth; heap-select when the depth
//   limit is exhausted; final insertion sort of <= 3 elements.
// Nodes are float4 {x, y, z, bits(index)}; the comparator looks at one coordinate.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define LB_KD_HD __host__ __device__ inline
#else
#define LB_KD_HD inline
struct alignas(16) float4
{
    float x, y, z, w;
};
#endif

namespace lb
{

LB_KD_HD float kd_key(const float4 &n, int axis)
{
    return axis == 0 ? n.x : (axis == 1 ? n.y : n.z);
}

LB_KD_HD void kd_swap(float4 *a, uint32_t i, uint32_t j)
{
    const float4 t = a[i];
    a[i] = a[j];
    a[j] = t;
}

LB_KD_HD uint32_t kd_floor_log2(uint32_t n)
{
    uint32_t l = 0;
    while (n > 1u)
    {
        n >>= 1;
        ++l;
    }
    return l;
}

// median of a[ia], a[ib], a[ic] swapped into a[result]
LB_KD_HD void kd_move_median_to_first(float4 *a, uint32_t result, uint32_t ia, uint32_t ib, uint32_t ic, int axis)
{
    const float ka = kd_key(a[ia], axis), kb = kd_key(a[ib], axis), kc = kd_key(a[ic], axis);
    uint32_t pick;
    if (ka < kb)
    {
        if (kb < kc)
            pick = ib;
        else if (ka < kc)
            pick = ic;
        else
            pick = ia;
    }
    else if (ka < kc)
        pick = ia;
    else if (kb < kc)
        pick = ic;
    else
        pick = ib;
    kd_swap(a, result, pick);
}

// Hoare partition around a[pivot] over [first, last); returns the cut
LB_KD_HD uint32_t kd_unguarded_partition(float4 *a, uint32_t first, uint32_t last, uint32_t pivot, int axis)
{
    const float kp = kd_key(a[pivot], axis);
    while (true)
    {
        while (kd_key(a[first], axis) < kp)
            ++first;
        --last;
        while (kp < kd_key(a[last], axis))
            --last;
        if (!(first < last))
            return first;
        kd_swap(a, first, last);
        ++first;
    }
}

LB_KD_HD void kd_push_heap(float4 *a, uint32_t first, int64_t hole, int64_t top, const float4 value, int axis)
{
    int64_t parent = (hole - 1) / 2;
    while (hole > top && kd_key(a[first + parent], axis) < kd_key(value, axis))
    {
        a[first + hole] = a[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[first + hole] = value;
}

LB_KD_HD void kd_adjust_heap(float4 *a, uint32_t first, int64_t hole, int64_t len, const float4 value, int axis)
{
    const int64_t top = hole;
    int64_t second = hole;
    while (second < (len - 1) / 2)
    {
        second = 2 * (second + 1);
        if (kd_key(a[first + second], axis) < kd_key(a[first + second - 1], axis))
            --second;
        a[first + hole] = a[first + second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2)
    {
        second = 2 * (second + 1);
        a[first + hole] = a[first + second - 1];
        hole = second - 1;
    }
    kd_push_heap(a, first, hole, top, value, axis);
}

// smallest (middle-first) elements of [first,last) end up heap-ordered in [first,middle)
LB_KD_HD void kd_heap_select(float4 *a, uint32_t first, uint32_t middle, uint32_t last, int axis)
{
    const int64_t len = static_cast<int64_t>(middle) - static_cast<int64_t>(first);
    if (len >= 2)
    {
        int64_t parent = (len - 2) / 2;
        while (true)
        {
            const float4 value = a[first + parent];
            kd_adjust_heap(a, first, parent, len, value, axis);
            if (parent == 0)
                break;
            --parent;
        }
    }
    for (uint32_t i = middle; i < last; ++i)
        if (kd_key(a[i], axis) < kd_key(a[first], axis))
        {
            const float4 value = a[i];
            a[i] = a[first];
            kd_adjust_heap(a, first, 0, len, value, axis);
        }
}

LB_KD_HD void kd_insertion_sort(float4 *a, uint32_t first, uint32_t last, int axis)
{
    if (first == last)
        return;
    for (uint32_t i = first + 1; i != last; ++i)
    {
        const float4 val = a[i];
        if (kd_key(val, axis) < kd_key(a[first], axis))
        {
            for (uint32_t k = i; k != first; --k)
                a[k] = a[k - 1];
            a[first] = val;
        }
        else
        {
            uint32_t hole = i;
            while (kd_key(val, axis) < kd_key(a[hole - 1], axis))
            {
                a[hole] = a[hole - 1];
                --hole;
            }
            a[hole] = val;
        }
    }
}

// The introselect loop from an arbitrary state (first, nth, last, depth_limit). The cooperative
// kernels run the first rounds in parallel and hand the narrowed state to this routine.
LB_KD_HD void kd_introselect_from(float4 *a, uint32_t first, uint32_t nth, uint32_t last, uint32_t depth_limit,
                                  int axis)
{
    while (last - first > 3u)
    {
        if (depth_limit == 0u)
        {
            kd_heap_select(a, first, nth + 1u, last, axis);
            kd_swap(a, first, nth);
            return;
        }
        --depth_limit;
        const uint32_t mid = first + (last - first) / 2u;
        kd_move_median_to_first(a, first, first + 1u, mid, last - 1u, axis);
        const uint32_t cut = kd_unguarded_partition(a, first + 1u, last, first, axis);
        if (cut <= nth)
            first = cut;
        else
            last = cut;
    }
    kd_insertion_sort(a, first, last, axis);
}

// std::nth_element(a+first, a+nth, a+last, key-on-axis less)
LB_KD_HD void kd_nth_element(float4 *a, uint32_t first, uint32_t nth, uint32_t last, int axis)
{
    if (first == last || nth == last)
        return;
    kd_introselect_from(a, first, nth, last, 2u * kd_floor_log2(last - first), axis);
}

// Range of the implicit tree node reached from [b,e) by following the low `depth` bits of `path`
// (most significant first; 0 = left child [b,mid), 1 = right child [mid+1,e)), mid = b + (e-b)/2
// (kdtree.hpp:200, 210-218). Returns false when the path leaves the tree.
LB_KD_HD bool kd_range_at(uint32_t m, uint32_t depth, uint32_t path, uint32_t *b_out, uint32_t *e_out)
{
    uint32_t b = 0u, e = m;
    for (uint32_t l = depth; l-- > 0u;)
    {
        if (b >= e)
            return false;
        const uint32_t mid = b + (e - b) / 2u;
        if ((path >> l) & 1u)
            b = mid + 1u;
        else
            e = mid;
    }
    *b_out = b;
    *e_out = e;
    return b < e;
}

// Pre-order rank of the node stored at array slot `slot` of an m-node implicit tree: node first,
// then the left subtree, then the right one (radius_search pushes right then left on a LIFO stack,
// kdtree.hpp:324-333).
LB_KD_HD uint32_t kd_preorder_rank_of_slot(uint32_t m, uint32_t slot)
{
    uint32_t b = 0u, e = m, rank = 0u;
    while (true)
    {
        const uint32_t mid = b + (e - b) / 2u;
        if (slot == mid)
            return rank;
        if (slot < mid)
        {
            rank += 1u;
            e = mid;
        }
        else
        {
            rank += 1u + (mid - b);
            b = mid + 1u;
        }
    }
}

} // namespace lb
