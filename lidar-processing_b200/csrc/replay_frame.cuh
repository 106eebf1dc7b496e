// Replay of the reference BFS (Clusterer::cluster, reference src/clustering.cpp:69-124), fourth generation (sm_100a):
// one WARP per r-connected component, the warps of a CTA share one frame's state bitmaps in shared memory.
//
// What the earlier generations taught (profiles/README.md, round 2): the replay is bound by instructions issued per
// expansion, not by bytes — a speculative CTA round cost ~8 500 warp instructions (window scan, per-entry candidate
// lists, CTA barriers, rank sort) to expand 2.3 entries, and every expansion loaded and tested several hundred
// candidates of which about 27 were still alive. This kernel removes both costs:
//
//  * state = two bitmaps per frame in SHARED memory indexed by `pos` (cell order: the points of a voxel cell are
//    contiguous): dead (removed) and que (already pushed). The candidates of an expansion are the set bits of ~dead inside
//    the bit ranges of the 27 neighbour cells, found by popc/ffs over a few words; only live points are loaded
//    (clustering.cpp:94-97 skips removed ones; most of a neighbourhood is removed long before it is visited again);
//  * the 27 neighbour cells come from one 108-byte row of the packed neighbour table the union-find built
//    (nb27[cell][k] = first pos | count << 20): one memory trip instead of 27 hash probes;
//  * an expansion is three dependent memory trips (entry record, neighbour row, live candidates) and a few hundred warp
//    instructions, with no CTA barrier: the warps of a CTA replay DIFFERENT components of the same frame and only share
//    the bitmaps (a component never writes another component's bits: r-components are closed under the radius test, so
//    a foreign live point that shows up in a neighbour cell just fails the distance test);
//  * a CTA serves one frame at a time (ticket = frame, several CTAs per frame each with its own bitmaps), warps claim
//    components from the frame's list, longest first. Components of one point are labelled without a replay.
//
// The FIFO of a component lives in global memory (queue[first member slot ...]); its head window of 64 entries is
// mirrored in the warp's shared memory. Pushes of one expansion enter the FIFO in ascending k-d pre-order rank — the
// order in which radius_search reports them (kdtree.hpp:292-341) — and only the first push of a point is kept (a later
// duplicate is a no-op when popped: the first pop always ends with the point removed).
#pragma once

#include "replay_cta3.cuh"

namespace lb
{

constexpr int kV4Warps = 8;
constexpr uint32_t kV4List = 256u;  // live candidates tested per pass (a pass covers up to 32 bitmap words)
constexpr uint32_t kV4Push = 256u;  // pushes of one expansion kept in shared memory (more spill to global memory)
constexpr uint32_t kV4Win = 64u;    // FIFO head window per warp
constexpr uint32_t kV4Buckets = 7u; // component size classes of the per-frame lists: >=4096, 1024, 256, 64, 16, 4, 2
constexpr uint32_t kV4MetaStride = 16u; // per-frame words: [0..6] counts, [7..13] fill cursors, [14] claim cursor
constexpr uint32_t kV4MaxPoints = 131072u; // frames above this do not fit two bitmaps beside the warps' buffers

struct __align__(16) V4Warp
{
    unsigned long long pbuf[kV4Push]; // pushes of the expansion: k-d rank << 32 | pos
    uint32_t list[kV4List];           // live candidates of the pass (pos)
    uint32_t hwin[kV4Win];            // pos of the FIFO entries [hbase, hbase + kV4Win)
};

struct __align__(16) V4Smem
{
    V4Warp w[kV4Warps];
    uint32_t ticket, pad[3];
};

LB_D uint32_t v4_bucket_of(uint32_t members)
{
    return members >= 4096u ? 0u : members >= 1024u ? 1u : members >= 256u ? 2u : members >= 64u ? 3u : members >= 16u ? 4u : members >= 4u ? 5u : 6u;
}

// ipts[pos] = {x, y, z, bits(k-d pre-order rank)}, seed_of[pos] = unset, member_pos[t] = pos of the t-th member of the
// component-sorted member list, cursor[0..15] = 0: everything the fourth-generation replay reads besides the union-find's
// tables. One thread per slot.
__global__ void __launch_bounds__(256)
replay_init4_kernel(const float4 *__restrict__ cpts, BatchView bv, const uint32_t *__restrict__ rank_of_point,
                    const uint32_t *__restrict__ member_idx, const uint32_t *__restrict__ pos_of, float4 *__restrict__ ipts,
                    uint32_t *__restrict__ seed_of, uint32_t *__restrict__ member_pos, uint32_t *__restrict__ cursor)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 16)
        cursor[threadIdx.x] = 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    {
        const float4 p = cpts[off + i];
        ipts[off + i] = make_float4(p.x, p.y, p.z, __uint_as_float(rank_of_point[off + __float_as_uint(p.w)]));
        seed_of[off + i] = kSeedUnset;
        member_pos[off + i] = pos_of[off + member_idx[off + i]];
    }
}

// Pass 1 over the member list (sorted by component, then index): counts the components of every frame per size class;
// a component of ONE point is finished here: its seed touches itself only (clustering.cpp:94-105), so it is a cluster
// of size 1 — valid or INVALID by the configured bounds.
__global__ void __launch_bounds__(256)
replay_complist_count_kernel(BatchView bv, CluParams prm, const uint32_t *__restrict__ member_root,
                             const uint32_t *__restrict__ member_idx, const uint32_t *__restrict__ member_pos,
                             const uint32_t *__restrict__ comp_size, uint32_t *__restrict__ seed_of,
                             uint8_t *__restrict__ seed_valid, uint32_t *__restrict__ frame_meta)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x)
    {
        const uint32_t r = member_root[off + t];
        if (t != 0u && member_root[off + t - 1u] == r)
            continue;
        const uint32_t size = comp_size[off + r];
        if (size == 1u)
        {
            const uint32_t idx = member_idx[off + t];
            seed_of[off + member_pos[off + t]] = idx;
            seed_valid[off + idx] = (1u < prm.min_cluster_size || 1u > prm.max_cluster_size) ? 0u : 1u;
        }
        else
            atomicAdd(&frame_meta[f * kV4MetaStride + v4_bucket_of(size)], 1u);
    }
}

// Pass 2: comp_list[off + slot] = first member slot of the component; the size classes are laid out longest first.
__global__ void __launch_bounds__(256)
replay_complist_fill_kernel(BatchView bv, const uint32_t *__restrict__ member_root, const uint32_t *__restrict__ comp_size,
                            uint32_t *__restrict__ frame_meta, uint32_t *__restrict__ comp_list)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    uint32_t *meta = frame_meta + f * kV4MetaStride;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x)
    {
        const uint32_t r = member_root[off + t];
        if (t != 0u && member_root[off + t - 1u] == r)
            continue;
        const uint32_t size = comp_size[off + r];
        if (size == 1u)
            continue;
        const uint32_t b = v4_bucket_of(size);
        uint32_t base = 0u;
        for (uint32_t k = 0; k < b; ++k)
            base += meta[k];
        comp_list[off + base + atomicAdd(&meta[kV4Buckets + b], 1u)] = t;
    }
}

// Sorts n 64-bit keys ascending with one warp; works on shared or global memory ("flip" bitonic network, see
// cta_bitonic_sort: slots past n behave like +infinity without being touched).
LB_D void warp_bitonic_sort(volatile unsigned long long *a, uint32_t n, uint32_t lane)
{
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
            __threadfence_block();
            __syncwarp();
        }
}

// Dynamic shared memory: V4Smem followed by the dead and que bitmaps of `bitmap_words` words each.
// tickets: ticket t serves frame t % frames; a frame is served by n_tickets / frames CTAs at most.
template <int MINB>
__global__ void __launch_bounds__(kV4Warps * 32, MINB)
replay_frame_kernel(const float4 *__restrict__ ipts_all, const uint32_t *__restrict__ cell_of_all,
                    const uint32_t *__restrict__ nb27_all, const uint2 *__restrict__ cinfo_all, BatchView bv, CluParams prm,
                    const uint32_t *__restrict__ member_root, const uint32_t *__restrict__ member_idx,
                    const uint32_t *__restrict__ member_pos, const uint32_t *__restrict__ comp_size,
                    uint32_t *__restrict__ seed_of, uint32_t *__restrict__ queue,
                    unsigned long long *__restrict__ push_spill, uint8_t *__restrict__ seed_valid,
                    const uint32_t *__restrict__ comp_list, uint32_t *__restrict__ frame_meta,
                    uint32_t *__restrict__ ticket_cursor, uint32_t n_tickets, uint32_t bitmap_words)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V4Smem &sm = *reinterpret_cast<V4Smem *>(smem_raw);
    uint32_t *dead = reinterpret_cast<uint32_t *>(smem_raw + sizeof(V4Smem));
    uint32_t *que = dead + bitmap_words;
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t lt = lanemask_lt();
    V4Warp &W = sm.w[tid >> 5];

    while (true)
    {
        __syncthreads();
        if (tid == 0)
        {
            const uint32_t t = atomicAdd(ticket_cursor, 1u);
            sm.ticket = t;
            // how far the frame's claim cursor is (other CTAs advance it: one reader, so that the CTA decides as one)
            sm.pad[0] = t < n_tickets ? *reinterpret_cast<volatile uint32_t *>(&frame_meta[(t % bv.frames) * kV4MetaStride + 2u * kV4Buckets]) : 0u;
        }
        __syncthreads();
        const uint32_t ticket = sm.ticket;
        if (ticket >= n_tickets)
            break;
        const uint32_t f = ticket % bv.frames;
        const uint32_t m = bv.cnt[f];
        const uint32_t off = bv.off[f];
        uint32_t *meta = frame_meta + f * kV4MetaStride;
        uint32_t n_comp = 0u;
        for (uint32_t k = 0; k < kV4Buckets; ++k)
            n_comp += meta[k];
        const uint32_t words = (m + 31u) >> 5;
        if (words > bitmap_words || sm.pad[0] >= n_comp)
            continue; // a frame that cannot be served here (never ticketed by the host), or nothing left to claim
        for (uint32_t i = tid; i < words; i += blockDim.x)
        {
            dead[i] = 0u;
            que[i] = 0u;
        }
        __syncthreads();

        const float4 *ip = ipts_all + off;
        const uint32_t *cof = cell_of_all + off;
        const uint32_t *nb = nb27_all + static_cast<size_t>(off) * 27u;
        const uint2 *ci = cinfo_all + off;
        uint32_t *so = seed_of + off;

        while (true) // ---- one component per iteration
        {
            uint32_t claim = 0u;
            if (lane == 0)
                claim = atomicAdd(&meta[2u * kV4Buckets], 1u);
            claim = __shfl_sync(kFullMask, claim, 0);
            if (claim >= n_comp)
                break;
            const uint32_t t_start = comp_list[off + claim];
            const uint32_t n_mem = comp_size[off + member_root[off + t_start]];
            uint32_t *qu = queue + off + t_start; // the component's FIFO (pos)
            unsigned long long *spill = push_spill + off + t_start;
            const uint32_t *midx = member_idx + off + t_start;
            const uint32_t *mpos = member_pos + off + t_start;

            uint32_t u = 0u; // next member (ascending index) to examine as a seed candidate (clustering.cpp:70-75)
            while (true)
            {
                uint32_t seed_l = 0xFFFFFFFFu;
                while (u < n_mem)
                {
                    const uint32_t uu = u + lane;
                    bool cand = false;
                    if (uu < n_mem)
                    {
                        const uint32_t p = __ldg(&mpos[uu]);
                        cand = ((dead[p >> 5] >> (p & 31u)) & 1u) == 0u;
                    }
                    const uint32_t bc = __ballot_sync(kFullMask, cand);
                    if (bc)
                    {
                        seed_l = u + static_cast<uint32_t>(__ffs(bc) - 1);
                        break;
                    }
                    u += 32u;
                }
                if (seed_l == 0xFFFFFFFFu)
                    break; // component done
                u = seed_l + 1u;
                const uint32_t seed_idx = __ldg(&midx[seed_l]);
                const uint32_t seed_pos = __ldg(&mpos[seed_l]);

                uint32_t head = 0u, tail = 1u, hbase = 0u, touched = 0u; // touched: this lane's share
                if (lane == 0)
                {
                    qu[0] = seed_pos;
                    W.hwin[0] = seed_pos;
                    atomicOr(&que[seed_pos >> 5], 1u << (seed_pos & 31u));
                }
                __syncwarp();

                while (head < tail) // clustering.cpp:80-111
                {
                    if (head >= hbase + 32u)
                    {
                        // the head left the lower half of the mirrored window: re-centre it (entries written by this warp)
                        hbase = head & ~31u;
                        __syncwarp();
                        for (uint32_t i = lane; i < kV4Win; i += 32u)
                        {
                            const uint32_t e = hbase + i;
                            if (e < tail)
                                W.hwin[e & (kV4Win - 1u)] = __ldcg(&qu[e]);
                        }
                        __syncwarp();
                    }
                    // ---- pop: the first live entry among the next 32; removed entries are no-ops (clustering.cpp:85-88)
                    uint32_t pos;
                    {
                        const uint32_t e = head + lane;
                        uint32_t p = 0u;
                        bool alive = false;
                        if (e < tail)
                        {
                            p = W.hwin[e & (kV4Win - 1u)];
                            alive = ((dead[p >> 5] >> (p & 31u)) & 1u) == 0u;
                        }
                        const uint32_t ba = __ballot_sync(kFullMask, alive);
                        if (ba == 0u)
                        {
                            head = min(tail, head + 32u);
                            continue;
                        }
                        const int first = __ffs(ba) - 1;
                        head += static_cast<uint32_t>(first) + 1u;
                        pos = __shfl_sync(kFullMask, p, first);
                    }
                    const float4 pj = __ldg(&ip[pos]);
                    const uint32_t cid = __ldg(&cof[pos]);

                    // ---- the 27 neighbour cells: bit ranges [start, start + count) of the bitmaps
                    uint32_t start = 0u, count = 0u;
                    if (lane < 27u)
                    {
                        const uint32_t v = __ldg(&nb[static_cast<size_t>(cid) * 27u + lane]);
                        start = v & kV3PosMask;
                        count = v >> kV3PosBits;
                        if (count == kV3CountSat)
                            count = __ldg(&ci[start]).x;
                    }
                    const uint32_t last = start + count - 1u;
                    const uint32_t nw = count ? (last >> 5) - (start >> 5) + 1u : 0u;
                    const uint32_t incl_w = warp_inclusive_scan(nw);
                    const uint32_t excl_w = incl_w - nw;
                    const uint32_t n_items = __shfl_sync(kFullMask, incl_w, 31);

                    uint32_t np = 0u; // pushes of this expansion
                    for (uint32_t item0 = 0u; item0 < n_items;)
                    {
                        // ---- a pass: the next (up to) 32 bitmap words, one per lane; their live bits are the candidates
                        const uint32_t i = item0 + lane;
                        uint32_t lo = 0u, hi = 26u; // first cell whose inclusive word prefix exceeds i
#pragma unroll
                        for (int it = 0; it < 5; ++it)
                        {
                            const uint32_t mid = (lo + hi) >> 1;
                            const uint32_t vv = __shfl_sync(kFullMask, incl_w, mid);
                            if (vv > i)
                                hi = mid;
                            else
                                lo = mid + 1u;
                        }
                        const uint32_t c = min(lo, 26u);
                        const uint32_t c_start = __shfl_sync(kFullMask, start, c), c_last = __shfl_sync(kFullMask, last, c),
                                       c_excl = __shfl_sync(kFullMask, excl_w, c);
                        uint32_t a = 0u, w = 0u;
                        if (i < n_items)
                        {
                            w = (c_start >> 5) + (i - c_excl);
                            a = ~dead[w];
                            if (w == (c_start >> 5))
                                a &= 0xFFFFFFFFu << (c_start & 31u);
                            if (w == (c_last >> 5))
                                a &= 0xFFFFFFFFu >> (31u - (c_last & 31u));
                        }
                        const uint32_t n = __popc(a);
                        const uint32_t incl = warp_inclusive_scan(n);
                        // the longest prefix of lanes whose candidates fit the list (a word holds at most 32: at least 8 lanes)
                        const uint32_t bf = __ballot_sync(kFullMask, incl <= kV4List);
                        const uint32_t n_lanes = bf == kFullMask ? 32u : static_cast<uint32_t>(__ffs(~bf) - 1);
                        const uint32_t tot = __shfl_sync(kFullMask, incl, n_lanes - 1u);
                        item0 += n_lanes;
                        if (tot == 0u)
                            continue;
                        if (lane < n_lanes)
                        {
                            uint32_t base = incl - n;
                            while (a)
                            {
                                const uint32_t b = __ffs(a) - 1u;
                                a &= a - 1u;
                                W.list[base++] = (w << 5) | b;
                            }
                        }
                        __syncwarp();
                        // ---- every live candidate is treated like the loop body of clustering.cpp:94-109
                        for (uint32_t g0 = 0u; g0 < tot; g0 += 64u)
                        {
                            uint32_t cpos2[2];
                            float4 cand2[2];
                            bool valid2[2];
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                            {
                                const uint32_t g = g0 + 32u * h + lane;
                                valid2[h] = g < tot;
                                cpos2[h] = 0u;
                                cand2[h] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (valid2[h])
                                {
                                    cpos2[h] = W.list[g];
                                    cand2[h] = __ldg(&ip[cpos2[h]]);
                                }
                            }
#pragma unroll
                            for (int h = 0; h < 2; ++h)
                            {
                                if (h == 1 && g0 + 32u >= tot)
                                    break;
                                const float4 cand = cand2[h];
                                const uint32_t cpos = cpos2[h];
                                bool push = false;
                                if (valid2[h])
                                {
                                    // KDTree::dist_sqr(target, node) (kdtree.hpp:145-163), inclusive test (kdtree.hpp:314)
                                    const float d2 = dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z);
                                    if (d2 <= prm.distance_squared)
                                    {
                                        ++touched; // indices_.push_back (with multiplicity)
                                        const uint32_t wbit = 1u << (cpos & 31u);
                                        if (d2 <= prm.inner_threshold)
                                        {
                                            so[cpos] = seed_idx; // clustering.cpp:99,102-105: leaves the cloud with this seed's label
                                            atomicOr(&dead[cpos >> 5], wbit);
                                        }
                                        else
                                            push = (atomicOr(&que[cpos >> 5], wbit) & wbit) == 0u; // clustering.cpp:106-109, first push only
                                    }
                                }
                                const uint32_t bp = __ballot_sync(kFullMask, push);
                                if (push)
                                {
                                    const uint32_t idx = np + __popc(bp & lt);
                                    const unsigned long long key =
                                        (static_cast<unsigned long long>(__float_as_uint(cand.w)) << 32) | static_cast<unsigned long long>(cpos);
                                    if (idx < kV4Push)
                                        W.pbuf[idx] = key;
                                    else
                                        spill[tail + idx] = key;
                                }
                                np += __popc(bp);
                            }
                        }
                        __syncwarp();
                    }

                    // ---- the FIFO receives this expansion's pushes in ascending k-d pre-order rank
                    if (np)
                    {
                        if (np <= 32u)
                        {
                            if (lane < np) // short lists: every key is ranked by counting the smaller ones
                            {
                                const unsigned long long key = W.pbuf[lane];
                                uint32_t dest = 0u;
                                for (uint32_t x = 0; x < np; ++x)
                                    dest += W.pbuf[x] < key ? 1u : 0u;
                                const uint32_t e = tail + dest;
                                const uint32_t p = static_cast<uint32_t>(key);
                                qu[e] = p;
                                if (e < hbase + kV4Win)
                                    W.hwin[e & (kV4Win - 1u)] = p;
                            }
                        }
                        else
                        {
                            volatile unsigned long long *pbuf = W.pbuf;
                            if (np > kV4Push)
                            {
                                // dense expansions: sort in global memory, the spill area holds entries kV4Push.. already
                                for (uint32_t i = lane; i < kV4Push; i += 32u)
                                    spill[tail + i] = W.pbuf[i];
                                pbuf = spill + tail;
                                __threadfence_block();
                                __syncwarp();
                            }
                            warp_bitonic_sort(pbuf, np, lane);
                            for (uint32_t i = lane; i < np; i += 32u)
                            {
                                const uint32_t e = tail + i;
                                const uint32_t p = static_cast<uint32_t>(pbuf[i]);
                                qu[e] = p;
                                if (e < hbase + kV4Win)
                                    W.hwin[e & (kV4Win - 1u)] = p;
                            }
                        }
                        tail += np;
                    }
                    __syncwarp();
                }
                // ---- seed finished: cluster size test with multiplicity (clustering.cpp:113-123)
                touched = warp_reduce_add(touched);
                if (lane == 0)
                    seed_valid[off + seed_idx] = (touched < prm.min_cluster_size || touched > prm.max_cluster_size) ? 0u : 1u;
            }
        }
    }
}

} // namespace lb
