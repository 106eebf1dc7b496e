// CTA-cooperative replay of the reference BFS for large r-connected components, third generation (sm_100a).
//
// Same speculative rounds as replay_cta.cuh (up to kCtaW live FIFO entries expanded against the state at the start of
// the round, sequential semantics restored in closed form; reference src/clustering.cpp:69-124). What changes is the
// amount of work and the number of dependent memory trips a round makes:
//
//  * the state of the replay is two bitmaps in SHARED memory indexed by `pos` (the frame's cell order, in which the
//    points of a voxel cell are contiguous): dead = removed or not a member of this component, que = already pushed.
//    The candidates of an expansion are therefore found by scanning the `~dead` bits of the 27 neighbour cells' bit
//    ranges: only points that are still ALIVE are ever loaded and tested. In the reference an expansion reports 27 live
//    neighbours on average while its 27 cells hold several hundred points, most of them removed long ago
//    (clustering.cpp:94-97 skips them one by one; the earlier generations loaded and tested every one of them);
//  * the 27 neighbour cells of an entry come from one 108-byte row of the packed neighbour table built for the
//    union-find (nb27[cell][k] = first pos | count << 20), not from 27 probes of the hash table: one memory trip;
//  * every warp derives the round's entries itself from the head window of the FIFO, which is mirrored in shared memory
//    (refilled every 256 consumed entries), so picking the entries costs no CTA barrier and no global load;
//  * the live candidates of ALL entries of the round go into one list in shared memory and are dealt over the whole CTA:
//    one trip to the point records per round, {x, y, z, k-d rank} in one 16-byte read-only load;
//  * nothing reads the bitmaps between the barrier that ends the candidate scan and the one that ends the tests, so
//    removals are written at once; a round has three CTA barriers.
//
// A round whose candidates do not fit the list is re-scanned for the longest prefix of entries that fits; a single
// entry with more live candidates than the list holds ("dense") is expanded alone, every thread testing the candidates
// of the bitmap words it scans, pushes spilling to global memory as in the earlier generations.
#pragma once

#include "replay_cta2.cuh"

namespace lb
{

constexpr uint32_t kV3ListCap = 4096u; // live candidates (hence pushes) of one round kept in shared memory
constexpr uint32_t kV3Win = 512u;      // FIFO head window mirrored in shared memory (two halves of 256)
constexpr uint32_t kV3PosBits = 20u;   // pos < 2^20: frames of at most 1 048 576 obstacle points
constexpr uint32_t kV3PosMask = (1u << kV3PosBits) - 1u;
constexpr uint32_t kV3CountSat = 4095u; // nb27 count field saturates here (the exact count is in cinfo then)
constexpr int kV3Unroll = 2;

LB_HD uint32_t nb27_pack(uint32_t start, uint32_t count)
{
    return count ? (start | ((count < kV3CountSat ? count : kV3CountSat) << kV3PosBits)) : 0u;
}

struct __align__(16) Cta3Smem
{
    unsigned long long pk[kV3ListCap]; // pushes of the round: (entry << 20 | k-d rank) << 32 | pos
    uint32_t list[kV3ListCap];         // live candidates of the round: entry << 20 | pos
    uint32_t hwin[kV3Win];             // pos of the FIFO entries [hbase, hbase + kV3Win)
    float4 ent[kCtaW];                 // entry coordinates
    uint32_t tk[kCtaW];                // live candidates of entry k
    uint32_t n_items, n_push, claim, found;
};

// ipts[pos] = {x, y, z, bits(k-d pre-order rank)}: the immutable point records of the third-generation replay
__global__ void __launch_bounds__(256)
replay_init3_kernel(const float4 *__restrict__ cpts, BatchView bv, const uint32_t *__restrict__ rank_of_point,
                    float4 *__restrict__ ipts)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < m; pos += gridDim.x * blockDim.x)
    {
        const float4 p = cpts[off + pos];
        ipts[off + pos] = make_float4(p.x, p.y, p.z, __uint_as_float(rank_of_point[off + __float_as_uint(p.w)]));
    }
}

// Job lists: buckets 0..3 = components (by size, longest first) of frames whose bitmaps fit the normal launch,
// bucket 4 = components of larger frames (1 CTA per SM, large bitmaps), bucket 5 = frames beyond that (first generation).
__global__ void __launch_bounds__(256)
replay_biglist3_kernel(BatchView bv, const uint32_t *__restrict__ member_root, const uint32_t *__restrict__ comp_size,
                       uint32_t cta_min_members, uint32_t normal_pts, uint32_t huge_pts, uint2 *__restrict__ biglist,
                       uint32_t bucket_capacity, uint32_t *__restrict__ big_count /* [0..3] */,
                       uint32_t *__restrict__ huge_count, uint32_t *__restrict__ legacy_count)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x)
    {
        const uint32_t r = member_root[off + t];
        if (t == 0u || member_root[off + t - 1u] != r)
        {
            const uint32_t size = comp_size[off + r];
            if (size >= cta_min_members)
            {
                if (m > huge_pts)
                    biglist[5u * bucket_capacity + atomicAdd(legacy_count, 1u)] = make_uint2(f, t);
                else if (m > normal_pts)
                    biglist[4u * bucket_capacity + atomicAdd(huge_count, 1u)] = make_uint2(f, t);
                else
                {
                    const uint32_t b = big_bucket_of(size);
                    biglist[b * bucket_capacity + atomicAdd(&big_count[b], 1u)] = make_uint2(f, t);
                }
            }
        }
    }
}

// Scans the live bits of the 27 neighbour cells of one entry (one whole warp) and appends entry << 20 | pos for every
// live candidate to the round's list. The unit of work is one 32-bit word of the bitmap: the words of the 27 bit ranges
// are dealt over the lanes (a dense cell near the sensor spans dozens of words, an average one a single word), each
// batch of 32 words reserves its list slots with one atomic. Returns the entry's number of live candidates (same value
// in every lane); slots past the list's capacity are counted, not written.
LB_D uint32_t v3_emit_entry(Cta3Smem &sm, const uint32_t *dead, const uint32_t *__restrict__ nb_row,
                            const uint2 *__restrict__ ci, uint32_t k, uint32_t lane)
{
    uint32_t start = 0u, count = 0u;
    if (lane < 27u)
    {
        const uint32_t v = __ldg(&nb_row[lane]);
        start = v & kV3PosMask;
        count = v >> kV3PosBits;
        if (count == kV3CountSat)
            count = __ldg(&ci[start]).x;
    }
    const uint32_t last = start + count - 1u;
    const uint32_t nw = count ? (last >> 5) - (start >> 5) + 1u : 0u;
    const uint32_t incl_w = warp_inclusive_scan(nw);
    const uint32_t excl_w = incl_w - nw;
    const uint32_t W = __shfl_sync(kFullMask, incl_w, 31);
    uint32_t T = 0u;
    for (uint32_t ib = 0; ib < W; ib += 32u)
    {
        const uint32_t i = ib + lane;
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
        if (i < W)
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
        const uint32_t tot = __shfl_sync(kFullMask, incl, 31);
        if (tot == 0u)
            continue;
        uint32_t base = 0u;
        if (lane == 0)
            base = atomicAdd(&sm.n_items, tot);
        base = __shfl_sync(kFullMask, base, 0) + incl - n;
        while (a)
        {
            const uint32_t b = __ffs(a) - 1u;
            a &= a - 1u;
            if (base < kV3ListCap)
                sm.list[base] = (k << kV3PosBits) | (w << 5) | b;
            ++base;
        }
        T += tot;
    }
    return T;
}

// MINB = CTAs per SM the register allocation is capped for. Dynamic shared memory: Cta3Smem followed by the two
// bitmaps of `bitmap_words` words each; the job lists only hold components of frames of at most 32 * bitmap_words points.
template <int MINB>
__global__ void __launch_bounds__(kCtaThreads, MINB)
replay_cta3_kernel(const float4 *__restrict__ ipts_all, const uint32_t *__restrict__ cell_of_all,
                   const uint32_t *__restrict__ nb27_all, const uint2 *__restrict__ cinfo_all, BatchView bv, CluParams prm,
                   const uint32_t *__restrict__ member_root, const uint32_t *__restrict__ member_idx,
                   const uint32_t *__restrict__ member_pos, const uint32_t *__restrict__ comp_size,
                   uint32_t *__restrict__ seed_of, uint32_t *__restrict__ queue,
                   unsigned long long *__restrict__ push_spill, uint8_t *__restrict__ seed_valid,
                   const uint2 *__restrict__ biglist, uint32_t bucket_capacity, const uint32_t *__restrict__ big_count,
                   uint32_t n_buckets, uint32_t *__restrict__ cursor, uint32_t bitmap_words,
                   uint32_t *__restrict__ job_stats /* optional: 8 words per job */)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cta3Smem &sm = *reinterpret_cast<Cta3Smem *>(smem_raw);
    uint32_t *dead = reinterpret_cast<uint32_t *>(smem_raw + sizeof(Cta3Smem));
    uint32_t *que = dead + bitmap_words;
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    uint32_t bucket_end[kBigBuckets];
    {
        uint32_t run = 0u;
        for (uint32_t b = 0; b < kBigBuckets; ++b)
        {
            run += b < n_buckets ? big_count[b] : 0u;
            bucket_end[b] = run;
        }
    }
    const uint32_t n_big = bucket_end[kBigBuckets - 1u];

    while (true)
    {
        __syncthreads();
        if (tid == 0)
            sm.claim = atomicAdd(cursor, 1u);
        __syncthreads();
        const uint32_t w = sm.claim;
        if (w >= n_big)
            break;
        uint32_t jb = 0u;
        while (w >= bucket_end[jb])
            ++jb;
        const uint2 job = biglist[jb * bucket_capacity + (w - (jb ? bucket_end[jb - 1u] : 0u))];
        const uint32_t f = job.x;
        const uint32_t t_start = job.y;
        const uint32_t m = bv.cnt[f];
        const uint32_t off = bv.off[f];
        const float4 *ip = ipts_all + off;
        const uint32_t *cof = cell_of_all + off;
        const uint32_t *nb = nb27_all + static_cast<size_t>(off) * 27u;
        const uint2 *ci = cinfo_all + off;
        uint32_t *so = seed_of + off;
        uint32_t *qu = queue + off + t_start; // the component's FIFO (pos)
        unsigned long long *spill = push_spill + off + t_start;
        const uint32_t *midx = member_idx + off + t_start;
        const uint32_t *mpos = member_pos + off + t_start;
        const uint32_t n_mem = comp_size[off + member_root[off + t_start]];
        const uint32_t words = (m + 31u) >> 5;
        if (words > bitmap_words)
            continue; // never listed (replay_biglist3_kernel routes by frame size); the labels would stay UNDEFINED

        for (uint32_t i = tid; i < words; i += kCtaThreads)
        {
            dead[i] = 0xFFFFFFFFu;
            que[i] = 0u;
        }
        if (tid == 0)
        {
            sm.n_items = 0u;
            sm.n_push = 0u;
        }
        __syncthreads();
        for (uint32_t t = tid; t < n_mem; t += kCtaThreads)
        {
            const uint32_t p = __ldg(&mpos[t]);
            atomicAnd(&dead[p >> 5], ~(1u << (p & 31u)));
        }
        __syncthreads();

        const long long job_t0 = clock64();
        uint32_t st_rounds = 0u, st_dense = 0u, st_cands = 0u, st_over = 0u;
        long long tA = 0, tB = 0, tC = 0, tmark = 0;
        uint32_t u = 0u; // next member (by ascending index) to examine as a seed candidate (clustering.cpp:70-75)
        while (true)
        {
            // ---- next seed: first member at or after u that is not removed
            uint32_t seed_l = 0xFFFFFFFFu;
            while (u < n_mem)
            {
                const uint32_t uu = u + tid;
                bool cand = false;
                if (uu < n_mem)
                {
                    const uint32_t p = __ldg(&mpos[uu]);
                    cand = ((dead[p >> 5] >> (p & 31u)) & 1u) == 0u;
                }
                const uint32_t bc = __ballot_sync(kFullMask, cand);
                if (tid == 0)
                    sm.found = 0xFFFFFFFFu;
                __syncthreads();
                if (bc && lane == 0)
                    atomicMin(&sm.found, u + warp * 32u + (__ffs(bc) - 1));
                __syncthreads();
                seed_l = sm.found;
                __syncthreads();
                if (seed_l != 0xFFFFFFFFu)
                    break;
                u += kCtaThreads;
            }
            if (seed_l == 0xFFFFFFFFu)
                break; // component done
            u = seed_l + 1u;
            const uint32_t seed_idx = midx[seed_l];
            const uint32_t seed_pos = mpos[seed_l];

            uint32_t head = 0u, tail = 1u, hbase = 0u, touched = 0u; // touched: this thread's share
            uint32_t take_cap = kCtaW; // entries taken per round: lowered when a round's candidates overflow the list
            if (tid == 0)
            {
                qu[0] = seed_pos;
                sm.hwin[0] = seed_pos;
                que[seed_pos >> 5] |= 1u << (seed_pos & 31u);
            }
            __syncthreads();

            while (head < tail) // clustering.cpp:80-111
            {
                tmark = clock64();
                if (head >= hbase + 256u)
                {
                    // the head left the lower half of the mirrored window: re-centre it
                    hbase = head & ~255u;
                    for (uint32_t i = tid; i < kV3Win; i += kCtaThreads)
                    {
                        const uint32_t e = hbase + i;
                        if (e < tail)
                            sm.hwin[e & (kV3Win - 1u)] = __ldcg(&qu[e]);
                    }
                    __syncthreads();
                }
                // ---- A: the first kCtaW live entries among the next 256 of the FIFO; every warp derives them itself
                // (lane i < n_take holds entry i). Removed entries are no-ops in the reference (clustering.cpp:85-88).
                uint32_t n_take = 0u, my_pos = 0u, my_widx = 0u;
                const uint32_t wend = min(tail, head + 256u);
                {
                    uint32_t wp[8], wb[8]; // the window in 8 coalesced slices: all shared-memory reads are issued together
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        const uint32_t e = head + 32u * i + lane;
                        uint32_t p = 0u;
                        bool alive = false;
                        if (e < wend)
                        {
                            p = sm.hwin[e & (kV3Win - 1u)];
                            alive = ((dead[p >> 5] >> (p & 31u)) & 1u) == 0u;
                        }
                        wp[i] = p;
                        wb[i] = __ballot_sync(kFullMask, alive);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                    {
                        const uint32_t cnt = __popc(wb[i]);
                        const uint32_t want = lane - n_take; // (wraps below n_take)
                        const bool takes = lane < take_cap && lane >= n_take && want < cnt;
                        uint32_t b = wb[i];
                        if (takes)
                            for (uint32_t x = 0; x < want; ++x)
                                b &= b - 1u;
                        const uint32_t src = takes ? static_cast<uint32_t>(__ffs(b) - 1) : 0u;
                        const uint32_t sp = __shfl_sync(kFullMask, wp[i], src);
                        if (takes)
                        {
                            my_pos = sp;
                            my_widx = 32u * i + src;
                        }
                        n_take = min(take_cap, n_take + cnt);
                    }
                }
                if (n_take == 0u)
                {
                    head = wend; // the whole window is dead
                    continue;
                }
                float4 pe = make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t my_cid = 0u;
                if (lane < n_take)
                {
                    pe = __ldg(&ip[my_pos]);
                    my_cid = __ldg(&cof[my_pos]);
                    if (warp == 0u)
                        sm.ent[lane] = pe; // read behind the next CTA barrier
                }
                ++st_rounds;

                // ---- which entries are really expanded: lane p < 28 tests the pair (j, k), j < k; every warp
                // derives the same mask. close bits of entry k sit at bit k(k-1)/2 + j.
                uint32_t applied = 0u;
                {
                    const uint32_t k = lane >= 21u ? 7u : lane >= 15u ? 6u : lane >= 10u ? 5u : lane >= 6u ? 4u : lane >= 3u ? 3u : lane >= 1u ? 2u : 1u;
                    const uint32_t j = (lane - ((k * (k - 1u)) >> 1)) & 7u;
                    const float ax = __shfl_sync(kFullMask, pe.x, j), ay = __shfl_sync(kFullMask, pe.y, j),
                                az = __shfl_sync(kFullMask, pe.z, j);
                    const float bx = __shfl_sync(kFullMask, pe.x, k), by = __shfl_sync(kFullMask, pe.y, k),
                                bz = __shfl_sync(kFullMask, pe.z, k);
                    const bool cl = lane < 28u && k < n_take && dist_sqr_ref(ax, ay, az, bx, by, bz) <= prm.inner_threshold;
                    const uint32_t pm = __ballot_sync(kFullMask, cl);
                    for (uint32_t kk = 0; kk < n_take; ++kk)
                    {
                        const uint32_t closebits = (pm >> ((kk * (kk - 1u)) >> 1)) & ((1u << kk) - 1u);
                        if ((closebits & applied) == 0u)
                            applied |= 1u << kk;
                    }
                }
                { const long long t = clock64(); tA += t - tmark; tmark = t; }

                // ---- B: warp k lists the live candidates of entry k (applied entries only)
                {
                    const bool mine = warp < n_take && ((applied >> warp) & 1u);
                    const uint32_t cid = __shfl_sync(kFullMask, my_cid, warp & 7u);
                    uint32_t T = 0u;
                    if (mine)
                        T = v3_emit_entry(sm, dead, nb + static_cast<size_t>(cid) * 27u, ci, warp, lane);
                    if (lane == 0)
                        sm.tk[warp] = T;
                }
                __syncthreads();
                uint32_t total = sm.n_items;
                uint32_t n_use = n_take;
                bool dense = false;
                if (total > kV3ListCap) // (uniform) rare: the round's candidates do not fit the list
                {
                    uint32_t run = 0u;
                    n_use = 0u;
                    for (uint32_t v = 0; v < n_take; ++v)
                    {
                        run += sm.tk[v];
                        if (run > kV3ListCap)
                            break;
                        n_use = v + 1u;
                    }
                    __syncthreads();
                    if (tid == 0)
                        sm.n_items = 0u;
                    __syncthreads();
                    if (n_use == 0u)
                    {
                        dense = true; // entry 0 alone has more live candidates than the list holds
                        n_use = 1u;
                        total = 0u;
                    }
                    else
                    {
                        const bool mine = warp < n_use && ((applied >> warp) & 1u);
                        const uint32_t cid = __shfl_sync(kFullMask, my_cid, warp & 7u);
                        if (mine)
                            v3_emit_entry(sm, dead, nb + static_cast<size_t>(cid) * 27u, ci, warp, lane);
                        __syncthreads();
                        total = sm.n_items;
                    }
                    applied &= (1u << n_use) - 1u;
                    take_cap = n_use;
                    ++st_over;
                }
                else if (2u * total <= kV3ListCap && take_cap < kCtaW)
                    ++take_cap;
                st_cands += total;
                { const long long t = clock64(); tB += t - tmark; tmark = t; }

                // ---- C: every live candidate of the round is treated like the loop body of clustering.cpp:94-109 by ONE
                // thread; what the entries expanded earlier in this round did to it follows from the geometry
                if (!dense)
                {
                    for (uint32_t base = 0; base < total; base += kCtaThreads * kV3Unroll)
                    {
                        uint32_t item2[kV3Unroll];
                        float4 cand2[kV3Unroll];
                        bool valid2[kV3Unroll];
#pragma unroll
                        for (int h = 0; h < kV3Unroll; ++h)
                        {
                            const uint32_t g = base + kCtaThreads * h + tid;
                            valid2[h] = g < total;
                            item2[h] = 0u;
                            cand2[h] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (valid2[h])
                            {
                                item2[h] = sm.list[g];
                                cand2[h] = __ldg(&ip[item2[h] & kV3PosMask]);
                            }
                        }
#pragma unroll
                        for (int h = 0; h < kV3Unroll; ++h)
                        {
                            if (!valid2[h])
                                continue;
                            const float4 cand = cand2[h];
                            const uint32_t pos = item2[h] & kV3PosMask;
                            const uint32_t k = item2[h] >> kV3PosBits;
                            const float4 pj = sm.ent[k];
                            // KDTree::dist_sqr(target, node) (kdtree.hpp:145-163), inclusive test (kdtree.hpp:314)
                            const float d2 = dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z);
                            if (!(d2 <= prm.distance_squared))
                                continue;
                            bool removed_before = false; // an earlier entry of the round removed the candidate
                            bool shared = false;         // an earlier entry of the round reaches the candidate
                            for (uint32_t em = applied & ((1u << k) - 1u); em; em &= em - 1u)
                            {
                                const float4 po = sm.ent[__ffs(em) - 1];
                                const float dj = dist_sqr_ref(po.x, po.y, po.z, cand.x, cand.y, cand.z);
                                removed_before |= dj <= prm.inner_threshold;
                                shared |= dj <= prm.distance_squared;
                            }
                            if (removed_before)
                                continue;
                            ++touched; // indices_.push_back (with multiplicity)
                            const uint32_t wbit = 1u << (pos & 31u);
                            if (d2 <= prm.inner_threshold)
                            {
                                so[pos] = seed_idx; // clustering.cpp:99,102-105: the point leaves the cloud with this seed's label
                                atomicOr(&dead[pos >> 5], wbit);
                            }
                            else if (!shared) // (an earlier entry that reaches it has pushed it, or it was queued before)
                            {
                                const uint32_t old = atomicOr(&que[pos >> 5], wbit); // clustering.cpp:106-109 (first push only)
                                if ((old & wbit) == 0u)
                                {
                                    const uint32_t idx = atomicAdd(&sm.n_push, 1u);
                                    sm.pk[idx] = (static_cast<unsigned long long>((k << kV3PosBits) | __float_as_uint(cand.w)) << 32) |
                                                 static_cast<unsigned long long>(pos);
                                }
                            }
                        }
                    }
                }
                else
                {
                    // dense entry, alone: warp w scans cells w, w + 8, ...; every lane tests the live candidates of the
                    // bitmap words it reads (each word has one reader, and only that reader changes its bits)
                    ++st_dense;
                    const float4 pj = sm.ent[0];
                    const uint32_t cid0 = __shfl_sync(kFullMask, my_cid, 0);
                    const uint32_t rowv = lane < 27u ? __ldg(&nb[static_cast<size_t>(cid0) * 27u + lane]) : 0u;
                    for (uint32_t c = warp; c < 27u; c += kCtaW)
                    {
                        const uint32_t v = __shfl_sync(kFullMask, rowv, c);
                        const uint32_t start = v & kV3PosMask;
                        uint32_t count = v >> kV3PosBits;
                        if (count == kV3CountSat)
                            count = __ldg(&ci[start]).x;
                        if (count == 0u)
                            continue;
                        const uint32_t last = start + count - 1u;
                        const uint32_t w0 = start >> 5, w1 = last >> 5;
                        for (uint32_t ww = w0 + lane; ww <= w1; ww += 32u)
                        {
                            uint32_t a = ~dead[ww];
                            if (ww == w0)
                                a &= 0xFFFFFFFFu << (start & 31u);
                            if (ww == w1)
                                a &= 0xFFFFFFFFu >> (31u - (last & 31u));
                            while (a)
                            {
                                const uint32_t b = __ffs(a) - 1u;
                                a &= a - 1u;
                                const uint32_t pos = (ww << 5) | b;
                                const float4 cand = __ldg(&ip[pos]);
                                const float d2 = dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z);
                                if (!(d2 <= prm.distance_squared))
                                    continue;
                                ++touched;
                                ++st_cands;
                                const uint32_t wbit = 1u << b;
                                if (d2 <= prm.inner_threshold)
                                {
                                    so[pos] = seed_idx;
                                    atomicOr(&dead[ww], wbit);
                                }
                                else
                                {
                                    const uint32_t old = atomicOr(&que[ww], wbit);
                                    if ((old & wbit) == 0u)
                                    {
                                        const uint32_t idx = atomicAdd(&sm.n_push, 1u);
                                        const unsigned long long key =
                                            (static_cast<unsigned long long>(__float_as_uint(cand.w)) << 32) |
                                            static_cast<unsigned long long>(pos);
                                        if (idx < kV3ListCap)
                                            sm.pk[idx] = key;
                                        else
                                            spill[tail + idx] = key;
                                    }
                                }
                            }
                        }
                    }
                }
                __syncthreads();
                { const long long t = clock64(); tC += t - tmark; tmark = t; }

                // ---- F: the FIFO receives the pushes ordered by (entry, k-d pre-order rank)
                const uint32_t np = sm.n_push;
                if (tid == 0)
                    sm.n_items = 0u;
                if (np)
                {
                    if (np <= kCtaThreads)
                    {
                        if (tid < np) // short lists: every key is ranked by counting the smaller ones
                        {
                            const unsigned long long key = sm.pk[tid];
                            uint32_t dest = 0u;
                            for (uint32_t x = 0; x < np; ++x)
                                dest += sm.pk[x] < key ? 1u : 0u;
                            const uint32_t e = tail + dest;
                            const uint32_t p = static_cast<uint32_t>(key) & kV3PosMask;
                            qu[e] = p;
                            if (e < hbase + kV3Win)
                                sm.hwin[e & (kV3Win - 1u)] = p;
                        }
                    }
                    else
                    {
                        volatile unsigned long long *pbuf = sm.pk;
                        if (np > kV3ListCap)
                        {
                            // dense rounds only: sort in global memory, the spill area holds entries kV3ListCap.. already
                            for (uint32_t i = tid; i < kV3ListCap; i += kCtaThreads)
                                spill[tail + i] = sm.pk[i];
                            pbuf = spill + tail;
                            __syncthreads();
                        }
                        cta_bitonic_sort(pbuf, np); // (uniform branch: np comes from shared memory)
                        for (uint32_t i = tid; i < np; i += kCtaThreads)
                        {
                            const uint32_t e = tail + i;
                            const uint32_t p = static_cast<uint32_t>(pbuf[i]) & kV3PosMask;
                            qu[e] = p;
                            if (e < hbase + kV3Win)
                                sm.hwin[e & (kV3Win - 1u)] = p;
                        }
                    }
                }
                head += __shfl_sync(kFullMask, my_widx, n_use - 1u) + 1u;
                tail += np;
                __syncthreads();
                if (tid == 0)
                    sm.n_push = 0u;
            }
            // ---- seed finished: cluster size test with multiplicity (clustering.cpp:113-123)
            touched = warp_reduce_add(touched);
            __syncthreads();
            if (lane == 0)
                sm.tk[warp] = touched;
            __syncthreads();
            if (tid == 0)
            {
                uint32_t tsum = 0u;
                for (uint32_t v = 0; v < kCtaW; ++v)
                    tsum += sm.tk[v];
                seed_valid[off + seed_idx] = (tsum < prm.min_cluster_size || tsum > prm.max_cluster_size) ? 0u : 1u;
            }
            __syncthreads();
        }
        if (job_stats && tid == 0)
        {
            uint32_t *js = job_stats + 8u * w;
            js[0] = f;
            js[1] = n_mem | (min(4095u, st_cands / max(1u, st_rounds)) << 20); // members | live candidates per round << 20
            js[2] = static_cast<uint32_t>((clock64() - job_t0) >> 10);
            js[3] = st_rounds;
            js[4] = st_dense | (st_over << 16); // dense rounds | rounds re-scanned for a shorter prefix << 16
            js[5] = static_cast<uint32_t>(tA >> 10);
            js[6] = static_cast<uint32_t>(tB >> 10);
            js[7] = static_cast<uint32_t>(tC >> 10);
        }
    }
}

} // namespace lb
