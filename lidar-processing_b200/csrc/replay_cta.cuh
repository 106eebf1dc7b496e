// CTA-cooperative replay of the reference BFS for LARGE r-connected components (sm_100a).
//
// The order-dependent BFS of Clusterer::cluster (reference src/clustering.cpp:69-124) is a sequential
// chain inside one component: a frame's latency is the longest chain (~1000 expansions), and a dense
// component scans thousands of candidates per expansion. One warp per component (cluster.cuh) pays
// five to six dependent L2 round trips per expansion; this kernel gives a whole CTA to a component
// and removes the chain in two ways:
//
//  * speculative rounds — the next kW live FIFO entries are expanded TOGETHER against the state at
//    the start of the round (window read, 27-cell lookups and candidate tests of all entries run in
//    parallel across the CTA, one L2 round trip each), then the sequential semantics are restored
//    in closed form, from geometry alone (all entries' coordinates sit in shared memory):
//        applied(k)      = entry k is not within the inner radius of an applied entry j < k
//                          (otherwise j removed it: the reference pops it and moves on)
//        removed_before  = some applied j < k holds the candidate within its inner radius
//        queued_before   = queued at round start, or some applied j < k holds it in its annulus
//    With these three predicates a candidate of entry k is treated exactly like the reference's
//    loop body at clustering.cpp:94-109 treats it; the round's pushes enter the FIFO ordered by
//    (entry, k-d pre-order rank), i.e. in the reference's order. Warp k owns entry k (lookups,
//    candidate tests, rank sort of its pushes); entries that are not applied cost nothing. A removal of
//    a candidate that an EARLIER entry of the round also reaches is only written behind the CTA barrier
//    that ends the candidate tests: the earlier entry touches it first in the reference's order and must
//    still see it alive, whichever warp runs first (else one count of the cluster size, which the
//    reference keeps with multiplicity, is lost). Every other state write is order-independent.
//  * direct rounds — an entry with more candidates than a warp's buffers hold (a dense
//    neighbourhood, thousands of candidates) is expanded alone by all 256 threads, its pushes are
//    sorted by a CTA-wide bitonic network in shared memory.
//
// The FIFO lives in global memory; its most recent kRing entries are mirrored in shared memory so a
// short queue (chain-like components) never waits for its own writes.
#pragma once

#include "cluster.cuh"

namespace lb
{

constexpr int kCtaThreads = 256;
constexpr uint32_t kCtaW = 8u;             // FIFO entries expanded per speculative round = warps per CTA
constexpr uint32_t kRing = 1024u;          // FIFO entries mirrored in shared memory (power of two)
constexpr uint32_t kEntryCandCap = 512u;   // candidates (hence pushes) of one entry in a speculative round
constexpr uint32_t kDirectPushCap = 4096u; // pushes of a direct round kept in shared memory
constexpr int kCtaUnroll = 4;

struct __align__(16) CtaSmem
{
    uint32_t ring[kRing];
    union
    {
        unsigned long long pbuf[kCtaW][kEntryCandCap]; // speculative round: pushes of entry k, rank << 31 | pos, from the front;
                                                       // postponed removals, slot << 32 | pos, from the back (a candidate is one or the other)
        unsigned long long dpush[kDirectPushCap];      // direct round: pushes of the single entry
    } u;
    uint8_t owner[kCtaW][kEntryCandCap]; // candidate number -> neighbour cell (0..26)
    float4 ent[kCtaW];                   // entry coordinates (w = state word)
    unsigned long long ent_key[kCtaW];   // cell key of the entry
    uint32_t ent_widx[kCtaW];            // window index of the entry
    uint32_t tk[kCtaW], np[kCtaW];       // candidates / pushes of entry k
    uint32_t dstart[27], dexcl[27], dincl[27], dslot[27]; // neighbour cells of entry 0 (direct round)
    uint32_t wcnt[8];
    uint32_t n_push, claim, found;
};
static_assert(sizeof(CtaSmem) <= 48u * 1024u, "static shared memory");

// Sorts n 64-bit keys ascending with the whole CTA; works on shared or global memory. Bitonic network
// in its "flip" form: every compare-exchange moves the smaller key to the lower index, so the slots
// past n behave like +infinity without ever being touched (no padding, any n).
template <int NT = kCtaThreads> LB_D void cta_bitonic_sort(volatile unsigned long long *a, uint32_t n)
{
    uint32_t n_pad = 2u;
    while (n_pad < n)
        n_pad <<= 1;
    for (uint32_t kk = 2u; kk <= n_pad; kk <<= 1)
        for (uint32_t jj = kk >> 1; jj > 0u; jj >>= 1)
        {
            for (uint32_t t = threadIdx.x; t < (n_pad >> 1); t += NT)
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
            __syncthreads();
        }
}

// Job list of the CTA path, bucketed by component size so that the longest replays start first
// (longest-processing-time-first): bucket b holds {frame, first member slot} of the components with
// at least kBigBucketMin[b] members; big_count[b] = entries in bucket b. Bucket b occupies
// biglist[b * bucket_capacity ...).
constexpr uint32_t kBigBuckets = 4u;
LB_D uint32_t big_bucket_of(uint32_t members)
{
    return members >= 8192u ? 0u : (members >= 2048u ? 1u : (members >= 768u ? 2u : 3u));
}

__global__ void __launch_bounds__(256)
replay_biglist_kernel(BatchView bv, const uint32_t *__restrict__ member_root, const uint32_t *__restrict__ comp_size,
                      uint32_t cta_min_members, uint2 *__restrict__ biglist, uint32_t bucket_capacity,
                      uint32_t *__restrict__ big_count)
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
                const uint32_t b = big_bucket_of(size);
                biglist[b * bucket_capacity + atomicAdd(&big_count[b], 1u)] = make_uint2(f, t);
            }
        }
    }
}

// MINB = CTAs per SM the register allocation is capped for (99 / 80 / 64 registers for 2 / 3 / 4): the kernel
// waits on CTA barriers most of the time, so more resident CTAs per SM hide each other's round latency.
template <int MINB>
__global__ void __launch_bounds__(kCtaThreads, MINB)
replay_cta_kernel(float4 *__restrict__ rpts_all, BatchView bv, TableView tv, const uint4 *__restrict__ cells,
                  CluParams prm, const uint32_t *__restrict__ member_root, const uint32_t *__restrict__ member_idx,
                  const uint32_t *__restrict__ member_pos, const uint32_t *__restrict__ comp_size,
                  const unsigned long long *__restrict__ pkey_all, uint32_t *__restrict__ tlive_all,
                  uint32_t *__restrict__ seed_of, uint32_t *__restrict__ queue,
                  unsigned long long *__restrict__ push_spill, uint8_t *__restrict__ seed_valid,
                  const uint2 *__restrict__ biglist, uint32_t bucket_capacity, const uint32_t *__restrict__ big_count,
                  uint32_t *__restrict__ cursor, uint32_t *__restrict__ job_stats /* optional: 8 words per job, see lidar_b200_last_replay_stats */)
{
    __shared__ CtaSmem sm;
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
    uint32_t bucket_end[kBigBuckets];
    {
        uint32_t run = 0u;
        for (uint32_t b = 0; b < kBigBuckets; ++b)
        {
            run += big_count[b];
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
        const uint32_t mask = table_mask(m, tv.tcap[f]);
        const uint4 *tab = cells + tv.toff[f];
        uint32_t *tlive = tlive_all + tv.toff[f];
        float4 *rp = rpts_all + off;
        const unsigned long long *pkey = pkey_all + off;
        uint32_t *stw = reinterpret_cast<uint32_t *>(rp) + 3; // state word = .w of the point record
        uint32_t *so = seed_of + off;
        uint32_t *qu = queue + off + t_start; // the component's FIFO
        unsigned long long *spill = push_spill + off + t_start;
        const uint32_t *mroot = member_root + off;
        const uint32_t *midx = member_idx + off;
        const uint32_t *mpos = member_pos + off;
        const uint32_t root = mroot[t_start];
        const uint32_t t_end = t_start + comp_size[off + root];

        const long long job_t0 = clock64();
        uint32_t st_rounds = 0u, st_direct = 0u, st_taken = 0u, st_seeds = 0u, st_cands = 0u;
        long long tA = 0, tBC = 0, tEF = 0, tmark = 0;
        uint32_t u = t_start; // next member to examine as a seed candidate (ascending index, clustering.cpp:70-75)
        while (true)
        {
            // ---- next seed: first member at or after u that is not removed
            uint32_t seed_t = 0xFFFFFFFFu;
            while (u < t_end)
            {
                const uint32_t uu = u + tid;
                bool cand = false;
                if (uu < t_end)
                    cand = (__ldcg(&stw[4u * mpos[uu]]) & kStRemoved) == 0u;
                const uint32_t bc = __ballot_sync(kFullMask, cand);
                if (tid == 0)
                    sm.found = 0xFFFFFFFFu;
                __syncthreads();
                if (bc && lane == 0)
                    atomicMin(&sm.found, u + warp * 32u + (__ffs(bc) - 1));
                __syncthreads();
                seed_t = sm.found;
                __syncthreads();
                if (seed_t != 0xFFFFFFFFu)
                    break;
                u += kCtaThreads;
            }
            if (seed_t == 0xFFFFFFFFu)
                break; // component done
            ++st_seeds;
            u = seed_t + 1u;
            const uint32_t seed_idx = midx[seed_t];
            const uint32_t seed_pos = mpos[seed_t];

            uint32_t head = 0u, tail = 1u, touched = 0u; // touched: this thread's share
            if (tid == 0)
            {
                qu[0] = seed_pos;
                sm.ring[0] = seed_pos;
                atomicOr(&stw[4u * seed_pos], kStQueued);
            }
            __syncthreads();

            while (head < tail) // clustering.cpp:80-111
            {
                // ---- A: window of the next 256 FIFO entries, the first kCtaW live ones are taken
                tmark = clock64();
                const uint32_t e = head + tid;
                float4 pe = make_float4(0.f, 0.f, 0.f, __uint_as_float(kStRemoved));
                unsigned long long pk = 0ull;
                if (e < tail)
                {
                    const uint32_t qpos = (tail - e <= kRing) ? sm.ring[e & (kRing - 1u)] : __ldcg(&qu[e]);
                    pe = __ldcg(&rp[qpos]);
                    pk = __ldg(&pkey[qpos]);
                }
                const bool alive = (__float_as_uint(pe.w) & kStRemoved) == 0u;
                const uint32_t ba = __ballot_sync(kFullMask, alive);
                if (lane == 0)
                    sm.wcnt[warp] = __popc(ba);
                __syncthreads();
                uint32_t before = 0u, total_alive = 0u;
#pragma unroll
                for (uint32_t v = 0; v < 8u; ++v)
                {
                    const uint32_t c = sm.wcnt[v];
                    before += v < warp ? c : 0u;
                    total_alive += c;
                }
                if (total_alive == 0u)
                {
                    head = min(tail, head + kCtaThreads);
                    __syncthreads();
                    continue;
                }
                const uint32_t arank = before + __popc(ba & lt);
                if (alive && arank < kCtaW)
                {
                    sm.ent[arank] = pe;
                    sm.ent_key[arank] = pk;
                    sm.ent_widx[arank] = tid;
                }
                const uint32_t n_take = min(kCtaW, total_alive);
                __syncthreads();
                ++st_rounds;
                st_taken += n_take;
                { const long long t = clock64(); tA += t - tmark; tmark = t; }

                // ---- which entries are really expanded: lane p < 28 tests the pair (j, k), j < k; every warp
                // derives the same mask. close bits of entry k sit at bit k(k-1)/2 + j.
                uint32_t applied = 0u;
                {
                    const uint32_t k = lane >= 21u ? 7u : lane >= 15u ? 6u : lane >= 10u ? 5u : lane >= 6u ? 4u : lane >= 3u ? 3u : lane >= 1u ? 2u : 1u;
                    const uint32_t j = lane - ((k * (k - 1u)) >> 1);
                    bool cl = false;
                    if (lane < 28u && k < n_take)
                    {
                        const float4 pa = sm.ent[j], pb = sm.ent[k];
                        cl = dist_sqr_ref(pa.x, pa.y, pa.z, pb.x, pb.y, pb.z) <= prm.inner_threshold;
                    }
                    const uint32_t pm = __ballot_sync(kFullMask, cl);
                    for (uint32_t kk = 0; kk < n_take; ++kk)
                    {
                        const uint32_t closebits = (pm >> ((kk * (kk - 1u)) >> 1)) & ((1u << kk) - 1u);
                        if ((closebits & applied) == 0u)
                            applied |= 1u << kk;
                    }
                }

                // ---- B: warp k looks up the 27 cells of entry k (applied entries only)
                const bool mine = warp < n_take && ((applied >> warp) & 1u);
                uint32_t start = 0u, count = 0u, slot = 0u, incl = 0u, excl = 0u, T = 0u;
                float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
                if (mine)
                {
                    pj = sm.ent[warp];
                    if (lane < 27u)
                    {
                        // neighbour key = own key + (dx, dy, dz) in the packed 21-bit fields (biased, no borrow)
                        const long long dk = static_cast<long long>(static_cast<int>(lane % 3u) - 1) +
                                             (static_cast<long long>(static_cast<int>((lane / 3u) % 3u) - 1) << 21) +
                                             (static_cast<long long>(static_cast<int>(lane / 9u) - 1) << 42);
                        cell_lookup_alive(tab, tlive, mask, sm.ent_key[warp] + static_cast<unsigned long long>(dk), &start,
                                          &count, &slot);
                    }
                    incl = warp_inclusive_scan(count);
                    excl = incl - count;
                    T = __shfl_sync(kFullMask, incl, 31);
                    st_cands += T;
                    if (warp == 0u && lane < 27u)
                    {
                        sm.dstart[lane] = start;
                        sm.dexcl[lane] = excl;
                        sm.dincl[lane] = incl;
                        sm.dslot[lane] = slot;
                    }
                }
                if (lane == 0)
                    sm.tk[warp] = T;
                __syncthreads();
                // entries from the first dense one on wait for a later round; a dense FIRST entry is expanded alone
                uint32_t n_use = n_take;
#pragma unroll
                for (uint32_t v = kCtaW; v-- > 0u;)
                    if (v < n_take && sm.tk[v] > kEntryCandCap)
                        n_use = v;
                { const long long t = clock64(); tBC += t - tmark; tmark = t; }
                uint32_t np_total = 0u;

                if (n_use != 0u)
                {
                    // ---- C: warp k treats the candidates of entry k like the loop body of clustering.cpp:94-109
                    uint32_t my_np = 0u;
                    uint32_t my_nd = 0u; // removals postponed to the write pass
                    if (mine && warp < n_use)
                    {
                        const uint32_t earlier = applied & ((1u << warp) - 1u);
                        uint8_t *own = sm.owner[warp];
                        for (uint32_t i = 0; i < count; ++i)
                            own[excl + i] = static_cast<uint8_t>(lane);
                        __syncwarp();
                        for (uint32_t base = 0; base < T; base += 32u * kCtaUnroll)
                        {
                            uint32_t pos2[kCtaUnroll], slot2[kCtaUnroll];
                            float4 cand2[kCtaUnroll];
#pragma unroll
                            for (int h = 0; h < kCtaUnroll; ++h)
                            {
                                const uint32_t q = base + 32u * h + lane;
                                const bool valid = q < T;
                                const uint32_t c = valid ? own[q] : 0u;
                                const uint32_t cstart = __shfl_sync(kFullMask, start, c);
                                const uint32_t cexcl = __shfl_sync(kFullMask, excl, c);
                                slot2[h] = __shfl_sync(kFullMask, slot, c);
                                pos2[h] = cstart + (q - cexcl);
                                cand2[h] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kStRemoved));
                                if (valid)
                                    cand2[h] = __ldcg(&rp[pos2[h]]);
                            }
#pragma unroll
                            for (int h = 0; h < kCtaUnroll; ++h)
                            {
                                if (h > 0 && base + 32u * h >= T)
                                    break;
                                const float4 cand = cand2[h];
                                const uint32_t sw = __float_as_uint(cand.w);
                                bool push = false, defer = false;
                                if ((sw & kStRemoved) == 0u) // removed points are skipped (clustering.cpp:94-97)
                                {
                                    // KDTree::dist_sqr(target, node) (kdtree.hpp:145-163), inclusive test (kdtree.hpp:314)
                                    const float d2 = dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z);
                                    if (d2 <= prm.distance_squared)
                                    {
                                        // what the entries expanded earlier in this round did to the candidate
                                        bool removed_before = false, queued_before = (sw & kStQueued) != 0u;
                                        bool shared = false; // an earlier entry of the round reaches the candidate too
                                        for (uint32_t em = earlier; em; em &= em - 1u)
                                        {
                                            const float4 po = sm.ent[__ffs(em) - 1];
                                            const float dj = dist_sqr_ref(po.x, po.y, po.z, cand.x, cand.y, cand.z);
                                            removed_before |= dj <= prm.inner_threshold;
                                            shared |= dj <= prm.distance_squared;
                                        }
                                        queued_before |= shared;
                                        if (!removed_before)
                                        {
                                            const uint32_t pos = pos2[h];
                                            so[pos] = seed_idx; // labels[k] = label (clustering.cpp:99)
                                            ++touched;          // indices_.push_back (with multiplicity)
                                            if (d2 <= prm.inner_threshold)
                                            {
                                                // clustering.cpp:102-105. An EARLIER entry that reaches this candidate
                                                // must still see it alive (it touches it first in the reference's
                                                // order), whichever warp gets here first: such a removal waits for
                                                // the write pass behind the CTA barrier. Later entries may see it at
                                                // once — the geometry tells them the same thing.
                                                if (shared)
                                                    defer = true;
                                                else
                                                {
                                                    atomicOr(&stw[4u * pos], kStRemoved);
                                                    atomicSub(&tlive[slot2[h]], 1u);
                                                }
                                            }
                                            else if (!queued_before)
                                            {
                                                atomicOr(&stw[4u * pos], kStQueued); // clustering.cpp:106-109 (first push only)
                                                push = true;
                                            }
                                        }
                                    }
                                }
                                const uint32_t bp = __ballot_sync(kFullMask, push);
                                if (push)
                                    sm.u.pbuf[warp][my_np + __popc(bp & lt)] =
                                        (static_cast<unsigned long long>(sw >> 2) << 31) | static_cast<unsigned long long>(pos2[h]);
                                my_np += __popc(bp);
                                const uint32_t bd = __ballot_sync(kFullMask, defer);
                                if (defer)
                                    sm.u.pbuf[warp][kEntryCandCap - 1u - (my_nd + __popc(bd & lt))] =
                                        (static_cast<unsigned long long>(slot2[h]) << 32) | static_cast<unsigned long long>(pos2[h]);
                                my_nd += __popc(bd);
                            }
                        }
                    }
                    if (lane == 0)
                        sm.np[warp] = my_np;
                    __syncthreads();
                    // ---- D: all state reads of the round are done; the postponed removals are written
                    for (uint32_t i = lane; i < my_nd; i += 32u)
                    {
                        const unsigned long long e = sm.u.pbuf[warp][kEntryCandCap - 1u - i];
                        atomicOr(&stw[4u * static_cast<uint32_t>(e)], kStRemoved);
                        atomicSub(&tlive[static_cast<uint32_t>(e >> 32)], 1u);
                    }
                    // ---- F: the FIFO receives the pushes ordered by (entry, k-d pre-order rank)
                    uint32_t pre = 0u;
#pragma unroll
                    for (uint32_t v = 0; v < kCtaW; ++v)
                    {
                        const uint32_t c = sm.np[v];
                        pre += v < warp ? c : 0u;
                        np_total += c;
                    }
                    for (uint32_t e2 = lane; e2 < my_np; e2 += 32u)
                    {
                        const unsigned long long key = sm.u.pbuf[warp][e2];
                        uint32_t dest = 0u;
                        for (uint32_t x = 0; x < my_np; ++x)
                            dest += sm.u.pbuf[warp][x] < key ? 1u : 0u;
                        const uint32_t pos = static_cast<uint32_t>(key) & 0x7FFFFFFFu;
                        qu[tail + pre + dest] = pos;
                        sm.ring[(tail + pre + dest) & (kRing - 1u)] = pos;
                    }
                    head += sm.ent_widx[n_use - 1u] + 1u;
                }
                else
                {
                    // ---- direct round: entry 0 alone, all 256 threads, acting on the loaded state at once
                    ++st_direct;
                    if (tid == 0)
                        sm.n_push = 0u;
                    __syncthreads();
                    pj = sm.ent[0];
                    T = sm.dincl[26];
                    for (uint32_t base = 0; base < T; base += kCtaThreads * kCtaUnroll)
                    {
                        uint32_t pos2[kCtaUnroll], ci2[kCtaUnroll];
                        float4 cand2[kCtaUnroll];
#pragma unroll
                        for (int h = 0; h < kCtaUnroll; ++h)
                        {
                            const uint32_t q = base + kCtaThreads * h + tid;
                            cand2[h] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kStRemoved));
                            pos2[h] = 0u;
                            ci2[h] = 0u;
                            if (q < T)
                            {
                                uint32_t lo = 0u, hi = 26u;
#pragma unroll
                                for (int it = 0; it < 5; ++it) // first cell whose inclusive prefix exceeds q
                                {
                                    const uint32_t mid = (lo + hi) >> 1;
                                    if (sm.dincl[mid] > q)
                                        hi = mid;
                                    else
                                        lo = mid + 1u;
                                }
                                ci2[h] = lo;
                                pos2[h] = sm.dstart[lo] + (q - sm.dexcl[lo]);
                                cand2[h] = __ldcg(&rp[pos2[h]]);
                            }
                        }
#pragma unroll
                        for (int h = 0; h < kCtaUnroll; ++h)
                        {
                            if (h > 0 && base + kCtaThreads * h >= T)
                                break;
                            const uint32_t pos = pos2[h];
                            const float4 cand = cand2[h];
                            const uint32_t sw = __float_as_uint(cand.w);
                            bool push = false;
                            if ((sw & kStRemoved) == 0u)
                            {
                                const float d2 = dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z);
                                if (d2 <= prm.distance_squared)
                                {
                                    so[pos] = seed_idx;
                                    ++touched;
                                    if (d2 <= prm.inner_threshold)
                                    {
                                        atomicOr(&stw[4u * pos], kStRemoved);
                                        atomicSub(&tlive[sm.dslot[ci2[h]]], 1u);
                                    }
                                    else if ((sw & kStQueued) == 0u)
                                    {
                                        atomicOr(&stw[4u * pos], kStQueued);
                                        push = true;
                                    }
                                }
                            }
                            const uint32_t bp = __ballot_sync(kFullMask, push);
                            if (bp)
                            {
                                uint32_t pb = 0u;
                                if (lane == 0)
                                    pb = atomicAdd(&sm.n_push, static_cast<uint32_t>(__popc(bp)));
                                pb = __shfl_sync(kFullMask, pb, 0);
                                if (push)
                                {
                                    const uint32_t idx = pb + __popc(bp & lt);
                                    const unsigned long long key =
                                        (static_cast<unsigned long long>(sw >> 2) << 31) | static_cast<unsigned long long>(pos);
                                    if (idx < kDirectPushCap)
                                        sm.u.dpush[idx] = key;
                                    else
                                        spill[tail + idx] = key;
                                }
                            }
                        }
                    }
                    __syncthreads();
                    np_total = sm.n_push;
                    unsigned long long *pbuf = sm.u.dpush;
                    if (np_total > kDirectPushCap)
                    {
                        // rare: sort in global memory (the spill area holds entries kDirectPushCap.. already)
                        for (uint32_t i = tid; i < kDirectPushCap; i += kCtaThreads)
                            spill[tail + i] = sm.u.dpush[i];
                        pbuf = spill + tail;
                        __syncthreads();
                    }
                    if (np_total != 0u)
                    {
                        if (np_total <= kCtaThreads)
                        {
                            if (tid < np_total)
                            {
                                const unsigned long long mine = pbuf[tid];
                                uint32_t dest = 0u;
                                for (uint32_t x = 0; x < np_total; ++x)
                                    dest += pbuf[x] < mine ? 1u : 0u;
                                const uint32_t pos = static_cast<uint32_t>(mine) & 0x7FFFFFFFu;
                                qu[tail + dest] = pos;
                                sm.ring[(tail + dest) & (kRing - 1u)] = pos;
                            }
                        }
                        else
                        {
                            cta_bitonic_sort(pbuf, np_total);
                            for (uint32_t i = tid; i < np_total; i += kCtaThreads)
                            {
                                const uint32_t pos = static_cast<uint32_t>(pbuf[i]) & 0x7FFFFFFFu;
                                qu[tail + i] = pos;
                                if (np_total - i <= kRing)
                                    sm.ring[(tail + i) & (kRing - 1u)] = pos;
                            }
                        }
                    }
                    head += sm.ent_widx[0] + 1u;
                }
                tail += np_total;
                __syncthreads();
                { const long long t = clock64(); tEF += t - tmark; tmark = t; }
            }
            // ---- seed finished: cluster size test with multiplicity (clustering.cpp:113-123)
            touched = warp_reduce_add(touched);
            __syncthreads();
            if (lane == 0)
                sm.wcnt[warp] = touched;
            __syncthreads();
            if (tid == 0)
            {
                uint32_t tsum = 0u;
                for (uint32_t v = 0; v < 8u; ++v)
                    tsum += sm.wcnt[v];
                seed_valid[off + seed_idx] = (tsum < prm.min_cluster_size || tsum > prm.max_cluster_size) ? 0u : 1u;
            }
            __syncthreads();
        }
        if (job_stats && tid == 0)
        {
            uint32_t *js = job_stats + 8u * w;
            js[0] = f;
            js[1] = t_end - t_start;
            js[2] = static_cast<uint32_t>((clock64() - job_t0) >> 10);
            js[3] = st_rounds;
            js[4] = st_direct;
            js[5] = static_cast<uint32_t>(tA >> 10);
            js[6] = static_cast<uint32_t>(tBC >> 10);
            js[7] = static_cast<uint32_t>(tEF >> 10);
        }
    }
}

} // namespace lb
