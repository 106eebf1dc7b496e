// Device emulation of the reference k-d tree build (reference src/kdtree.hpp:174-225): one
// std::nth_element per tree node, re-enacted so that the node permutation — and hence the pre-order
// rank of every point — equals libstdc++'s bit for bit (see kd_select.h for why that matters).
//
// Level-synchronous: all nodes of one depth are independent. Large ranges run the introselect
// rounds cooperatively (one CTA or one warp per range): the Hoare partition of a round is evaluated
// in closed form — the k-th "not < pivot" element from the left swaps with the k-th "not > pivot"
// element from the right while the former lies left of the latter — using ballot ranks and one
// cross-warp scan, which produces exactly the sequential algorithm's permutation and cut. Narrowed
// ranges and the bottom of the tree are finished by single threads running the sequential
// transcription (kd_select.h).
#pragma once

#include "common.cuh"
#include "kd_select.h"

#include <cstdlib>

namespace lb
{

constexpr uint32_t kKdSeqCutoff = 48u;    // cooperative rounds stop once a working range is this small
constexpr uint32_t kKdSubtreeMax = 64u;   // ranges of at most this many nodes are finished by one thread

// nodes[i] = {x, y, z, bits(i)} in input order (kdtree.hpp:186-190)
__global__ void __launch_bounds__(256)
kd_init_kernel(const float4 *__restrict__ pts, BatchView bv, float4 *__restrict__ nodes)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    {
        const float4 p = pts[off + i];
        nodes[off + i] = make_float4(p.x, p.y, p.z, __uint_as_float(i));
    }
}

template <int NT> struct KdCoopSmem
{
    uint32_t ge_w[32];
    uint32_t le_w[32];
    uint32_t first, last, depth_limit, done;
    uint32_t K, min_unswapped_ge, min_swapped_le;
};

// barrier of a group of NT threads: a warp, or NT/32 whole warps meeting at the named barrier `bar_id` (0 = the
// barrier __syncthreads() uses: right when the group is the whole CTA)
template <int NT> LB_D void kd_block_sync(uint32_t bar_id)
{
    if (NT > 32)
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NT) : "memory");
    else
        __syncwarp();
}

// std::nth_element(nodes+first, nodes+nth, nodes+last) by a group of NT threads (`tid` = thread number inside the
// group, `bar_id` = the group's named barrier). nodes / gepos / lepos may live in global or in shared memory.
template <int NT>
LB_D void kd_coop_nth_element(float4 *nodes, uint32_t first, uint32_t nth, uint32_t last, int axis, uint32_t *gepos,
                              uint32_t *lepos, KdCoopSmem<NT> &sm, uint32_t tid = threadIdx.x, uint32_t bar_id = 0u)
{
    constexpr uint32_t NW = NT / 32;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31u;
    if (first == last || nth == last)
        return;
    uint32_t depth_limit = 2u * kd_floor_log2(last - first);

    while (last - first > 3u && last - first > kKdSeqCutoff && depth_limit != 0u)
    {
        --depth_limit;
        if (tid == 0)
        {
            const uint32_t mid = first + (last - first) / 2u;
            kd_move_median_to_first(nodes, first, first + 1u, mid, last - 1u, axis);
            sm.K = 0u;
            sm.min_unswapped_ge = 0xFFFFFFFFu;
            sm.min_swapped_le = 0xFFFFFFFFu;
        }
        kd_block_sync<NT>(bar_id);
        const float kp = kd_key(nodes[first], axis);
        const uint32_t rb = first + 1u;
        const uint32_t r = last - rb;
        uint32_t chunk = (r + NW - 1u) / NW;
        chunk = (chunk + 31u) / 32u * 32u;
        const uint32_t cb = rb + warp * chunk;
        const uint32_t ce = min(last, cb + chunk);

        // pass 1: per-warp counts of "not < pivot" (GE) and "not > pivot" (LE)
        uint32_t ge_cnt = 0u, le_cnt = 0u;
        if (cb < last)
            for (uint32_t itb = cb; itb < ce; itb += 32u)
            {
                const uint32_t p = itb + lane;
                bool is_ge = false, is_le = false;
                if (p < ce)
                {
                    const float k = kd_key(nodes[p], axis);
                    is_ge = !(k < kp);
                    is_le = !(kp < k);
                }
                ge_cnt += __popc(__ballot_sync(kFullMask, is_ge));
                le_cnt += __popc(__ballot_sync(kFullMask, is_le));
            }
        if (lane == 0)
        {
            sm.ge_w[warp] = ge_cnt;
            sm.le_w[warp] = le_cnt;
        }
        kd_block_sync<NT>(bar_id);
        uint32_t ge_before = 0u, le_after = 0u;
        for (uint32_t v = 0; v < NW; ++v)
        {
            if (v < warp)
                ge_before += sm.ge_w[v];
            if (v > warp)
                le_after += sm.le_w[v];
        }

        // pass 2: closed-form swap partners
        uint32_t run_ge = 0u, run_le = 0u, k_local = 0u;
        uint32_t my_min_unswapped = 0xFFFFFFFFu, my_min_swapped_le = 0xFFFFFFFFu;
        if (cb < last)
            for (uint32_t itb = cb; itb < ce; itb += 32u)
            {
                const uint32_t p = itb + lane;
                bool is_ge = false, is_le = false;
                if (p < ce)
                {
                    const float k = kd_key(nodes[p], axis);
                    is_ge = !(k < kp);
                    is_le = !(kp < k);
                }
                const uint32_t bge = __ballot_sync(kFullMask, is_ge);
                const uint32_t ble = __ballot_sync(kFullMask, is_le);
                const uint32_t lt = lanemask_lt();
                const uint32_t ge_left = ge_before + run_ge + __popc(bge & lt);
                const uint32_t le_right = le_after + (le_cnt - run_le - __popc(ble & (lt | (1u << lane))));
                if (is_ge)
                {
                    if (le_right > ge_left)
                    {
                        gepos[rb + ge_left] = p;
                        ++k_local;
                    }
                    else
                        my_min_unswapped = min(my_min_unswapped, p);
                }
                if (is_le && ge_left > le_right)
                {
                    lepos[rb + le_right] = p;
                    my_min_swapped_le = min(my_min_swapped_le, p);
                }
                run_ge += __popc(bge);
                run_le += __popc(ble);
            }
        k_local = warp_reduce_add(k_local);
        my_min_unswapped = warp_reduce_min(my_min_unswapped);
        my_min_swapped_le = warp_reduce_min(my_min_swapped_le);
        if (lane == 0)
        {
            if (k_local)
                atomicAdd(&sm.K, k_local);
            atomicMin(&sm.min_unswapped_ge, my_min_unswapped);
            atomicMin(&sm.min_swapped_le, my_min_swapped_le);
        }
        __threadfence_block();
        kd_block_sync<NT>(bar_id);
        const uint32_t K = sm.K;
        for (uint32_t k = tid; k < K; k += NT)
            kd_swap(nodes, gepos[rb + k], lepos[rb + k]);
        const uint32_t cut = min(sm.min_unswapped_ge, sm.min_swapped_le);
        if (cut <= nth)
            first = cut;
        else
            last = cut;
        __threadfence_block();
        kd_block_sync<NT>(bar_id);
    }
    if (tid == 0)
        kd_introselect_from(nodes, first, nth, last, depth_limit, axis);
    __threadfence_block();
    kd_block_sync<NT>(bar_id);
}

// One CTA of NT threads per tree node of depth `depth`. grid = (2^depth, frames).
template <int NT>
__global__ void __launch_bounds__(NT)
kd_level_kernel(float4 *__restrict__ nodes, BatchView bv, uint32_t depth, uint32_t *__restrict__ gepos,
                uint32_t *__restrict__ lepos)
{
    __shared__ KdCoopSmem<NT> sm;
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    uint32_t b, e;
    if (!kd_range_at(m, depth, blockIdx.x, &b, &e))
        return;
    const uint32_t mid = b + (e - b) / 2u;
    kd_coop_nth_element<NT>(nodes + off, b, mid, e, static_cast<int>(depth % 3u), gepos + off, lepos + off, sm);
}

// One thread per tree node of depth `depth`; the thread finishes the node's whole subtree.
__global__ void __launch_bounds__(128)
kd_subtree_kernel(float4 *__restrict__ nodes, BatchView bv, uint32_t depth)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t path = blockIdx.x * blockDim.x + threadIdx.x;
    if (path >= (1u << depth))
        return;
    uint32_t b, e;
    if (!kd_range_at(m, depth, path, &b, &e))
        return;
    float4 *a = nodes + off;
    // explicit stack: ranges halve, so the depth below here is bounded by log2(range)+1
    uint32_t sb[40], se[40], sd[40];
    int top = 0;
    sb[0] = b;
    se[0] = e;
    sd[0] = depth;
    top = 1;
    while (top > 0)
    {
        --top;
        const uint32_t rb = sb[top], re = se[top], rd = sd[top];
        if (rb >= re)
            continue;
        const uint32_t mid = rb + (re - rb) / 2u;
        if (re - rb > 1u)
            kd_nth_element(a, rb, mid, re, static_cast<int>(rd % 3u));
        if (mid > rb && top < 39)
        {
            sb[top] = rb;
            se[top] = mid;
            sd[top] = rd + 1u;
            ++top;
        }
        if (mid + 1u < re && top < 39)
        {
            sb[top] = mid + 1u;
            se[top] = re;
            sd[top] = rd + 1u;
            ++top;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The bottom of the tree in SHARED memory: one CTA of 256 threads takes the tree node of depth `depth` whose range
// holds at most kKdFusedMax nodes, loads the range once and re-enacts EVERY remaining level there - the same
// cooperative introselect rounds, run by 256 / 128 / 64 / 32 threads per node as the nodes halve (named barriers per
// group), then one warp per node, then one thread per subtree of at most kKdSubtreeMax nodes - and writes the final
// permutation back. The level-synchronous version paid a launch, a drain and several L2 round trips per level and
// per introselect round for ranges that fit a fraction of one SM's shared memory.
constexpr uint32_t kKdFusedMax = 2048u;
constexpr int kKdFusedThreads = 256;

struct KdFusedSmem
{
    float4 nodes[kKdFusedMax];
    uint32_t gepos[kKdFusedMax];
    uint32_t lepos[kKdFusedMax];
    KdCoopSmem<32> grp[kKdFusedThreads / 32]; // per-group scratch (the layout does not depend on NT)
};

template <int NT>
LB_D void kd_fused_level(KdFusedSmem &sm, uint32_t len, uint32_t level, int axis, uint32_t n_ranges)
{
    constexpr uint32_t kGroups = kKdFusedThreads / NT;
    const uint32_t g = threadIdx.x / NT, gtid = threadIdx.x % NT;
    for (uint32_t r = g; r < n_ranges; r += kGroups)
    {
        uint32_t b, e;
        if (!kd_range_at(len, level, r, &b, &e) || e - b < 2u)
            continue; // (uniform for the whole group)
        kd_coop_nth_element<NT>(sm.nodes, b, b + (e - b) / 2u, e, axis, sm.gepos, sm.lepos,
                                reinterpret_cast<KdCoopSmem<NT> &>(sm.grp[g]), gtid, 1u + g);
    }
}

// fixed_levels = kKdFusedAll: every remaining level, down to the subtrees of at most kKdSubtreeMax nodes, which are
// finished here too (one frame in flight: the shortest dependent chain). Otherwise exactly that many levels for every
// range, then the permutation goes back to global memory and kd_subtree_kernel finishes it at full occupancy (batches:
// the thread-per-subtree tail would keep one warp of every 56 KB CTA busy while the union-find wants the SMs).
constexpr uint32_t kKdFusedAll = 0xFFFFFFFFu;

__global__ void __launch_bounds__(kKdFusedThreads)
kd_fused_kernel(float4 *__restrict__ nodes, BatchView bv, uint32_t depth, uint32_t fixed_levels)
{
    extern __shared__ __align__(16) unsigned char kd_fused_raw[];
    KdFusedSmem &sm = *reinterpret_cast<KdFusedSmem *>(kd_fused_raw);
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    uint32_t b0, e0;
    if (!kd_range_at(m, depth, blockIdx.x, &b0, &e0))
        return;
    const uint32_t len = e0 - b0; // <= kKdFusedMax by the host's level schedule
    if (len > kKdFusedMax)
        return;
    float4 *a = nodes + off + b0;
    for (uint32_t i = threadIdx.x; i < len; i += kKdFusedThreads)
        sm.nodes[i] = a[i];
    __syncthreads();
    uint32_t level = 0u;
    for (;; ++level)
    {
        const uint32_t n_ranges = 1u << level;
        if (fixed_levels == kKdFusedAll ? ((len + n_ranges - 1u) >> level) <= kKdSubtreeMax : level >= fixed_levels)
            break;
        const int axis = static_cast<int>((depth + level) % 3u);
        if (level == 0u)
            kd_fused_level<256>(sm, len, level, axis, n_ranges);
        else if (level == 1u)
            kd_fused_level<128>(sm, len, level, axis, n_ranges);
        else if (level == 2u)
            kd_fused_level<64>(sm, len, level, axis, n_ranges);
        else
            kd_fused_level<32>(sm, len, level, axis, n_ranges);
        __syncthreads();
    }
    // one thread per remaining subtree (ranges of at most kKdSubtreeMax nodes), sequential transcription in shared memory
    for (uint32_t path = threadIdx.x; fixed_levels == kKdFusedAll && path < (1u << level); path += kKdFusedThreads)
    {
        uint32_t b, e;
        if (!kd_range_at(len, level, path, &b, &e))
            continue;
        uint32_t sb[24], se[24], sd[24];
        int top = 1;
        sb[0] = b;
        se[0] = e;
        sd[0] = depth + level;
        while (top > 0)
        {
            --top;
            const uint32_t rb = sb[top], re = se[top], rd = sd[top];
            if (rb >= re)
                continue;
            const uint32_t mid = rb + (re - rb) / 2u;
            if (re - rb > 1u)
                kd_nth_element(sm.nodes, rb, mid, re, static_cast<int>(rd % 3u));
            if (mid > rb && top < 23)
            {
                sb[top] = rb;
                se[top] = mid;
                sd[top] = rd + 1u;
                ++top;
            }
            if (mid + 1u < re && top < 23)
            {
                sb[top] = mid + 1u;
                se[top] = re;
                sd[top] = rd + 1u;
                ++top;
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < len; i += kKdFusedThreads)
        a[i] = sm.nodes[i];
}

// rank_of_point[index] = pre-order rank of the node holding that point
__global__ void __launch_bounds__(256)
kd_rank_kernel(const float4 *__restrict__ nodes, BatchView bv, uint32_t *__restrict__ rank_of_point)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < m; s += gridDim.x * blockDim.x)
    {
        const uint32_t idx = __float_as_uint(nodes[off + s].w);
        rank_of_point[off + idx] = kd_preorder_rank_of_slot(m, s);
    }
}

// Host-side level schedule. max_m bounds every frame's point count. Returns kernel launches issued.
inline int kd_build_launch(cudaStream_t stream, const float4 *pts, BatchView bv, uint32_t max_m, float4 *nodes,
                           uint32_t *gepos, uint32_t *lepos, uint32_t *rank_of_point)
{
    if (max_m == 0u || bv.frames == 0u)
        return 0;
    int launches = 0;
    const uint32_t gx = (max_m + 255u) / 256u;
    kd_init_kernel<<<dim3(gx, bv.frames), 256, 0, stream>>>(pts, bv, nodes);
    ++launches;
    uint32_t depth = 0u;
    // size of the largest range at `depth` is at most ceil(max_m / 2^depth)
    // The fused bottom levels shorten the dependent chain of ONE frame (fewer launches and drains: the latency path);
    // in a large batch the level-synchronous kernels fill the machine anyway and the long-lived 56 KB CTAs of the
    // fused kernel only take SMs away from the union-find running beside them (measured: union_find 3.2 -> 5.9 ms per
    // 154 frames). LIDAR_B200_KD_FUSED=0/1 forces either.
    // LIDAR_B200_KD_FUSED=2: the levels between 2048 and 64 nodes per range in shared memory, subtrees by kd_subtree_kernel
    // (measured on 154-frame batches: 10 721 against 11 257 frames/s - the 56 KB CTAs of the shared-memory levels cost the
    // union-find beside them more than the five level launches they replace; not the default).
    static const int fused_env = std::getenv("LIDAR_B200_KD_FUSED") ? std::atoi(std::getenv("LIDAR_B200_KD_FUSED")) : -1;
    const bool fused = fused_env < 0 ? bv.frames <= 4u : fused_env != 0;
    const bool fused_middle = fused_env == 2;
    while (((max_m + (1u << depth) - 1u) >> depth) > (fused ? kKdFusedMax : kKdSubtreeMax))
    {
        const uint32_t range = (max_m + (1u << depth) - 1u) >> depth;
        const dim3 grid(1u << depth, bv.frames);
        if (range >= 8192u)
            kd_level_kernel<1024><<<grid, 1024, 0, stream>>>(nodes, bv, depth, gepos, lepos);
        else if (range >= 1024u)
            kd_level_kernel<256><<<grid, 256, 0, stream>>>(nodes, bv, depth, gepos, lepos);
        else
            kd_level_kernel<32><<<grid, 32, 0, stream>>>(nodes, bv, depth, gepos, lepos);
        ++launches;
        ++depth;
    }
    if (fused)
    {
        static bool attr_done = false; // (a per-device function attribute; contexts of one process share the device)
        if (!attr_done)
        {
            cudaFuncSetAttribute(kd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(KdFusedSmem)));
            attr_done = true;
        }
        uint32_t levels = kKdFusedAll;
        if (fused_middle)
        {
            levels = 0u;
            while (((max_m + (1u << (depth + levels)) - 1u) >> (depth + levels)) > kKdSubtreeMax)
                ++levels;
        }
        kd_fused_kernel<<<dim3(1u << depth, bv.frames), kKdFusedThreads, sizeof(KdFusedSmem), stream>>>(nodes, bv, depth, levels);
        ++launches;
        if (fused_middle)
        {
            depth += levels;
            const uint32_t paths = 1u << depth;
            kd_subtree_kernel<<<dim3((paths + 127u) / 128u, bv.frames), 128, 0, stream>>>(nodes, bv, depth);
            ++launches;
        }
    }
    else
    {
        const uint32_t paths = 1u << depth;
        kd_subtree_kernel<<<dim3((paths + 127u) / 128u, bv.frames), 128, 0, stream>>>(nodes, bv, depth);
        ++launches;
    }
    kd_rank_kernel<<<dim3(gx, bv.frames), 256, 0, stream>>>(nodes, bv, rank_of_point);
    ++launches;
    return launches;
}

} // namespace lb
