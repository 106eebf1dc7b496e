// CTA-cooperative replay of the reference BFS for large r-connected components, second generation (sm_100a).
//
// Same algorithm and the same round structure as replay_cta.cuh (speculative rounds of up to kCtaW FIFO entries whose
// sequential semantics are restored in closed form, direct rounds for dense entries; reference src/clustering.cpp:69-124)
// — what changes is WHERE the data of a round lives, because the first generation spent its time waiting for L2:
//
//  * the component's mutable state (removed / queued, 2 bits per member) is a bitmap in SHARED memory, indexed by the
//    member's local id (lid = rank of the point index inside the component). Nothing a round reads from global memory
//    is ever written during the replay any more, so
//  * every point record is immutable and read through the read-only path (ld.global.nc, L1-cached): candidates from
//    ipts[pos] = {x, y, z, member slot} (cell order: the points of a cell are contiguous) and their k-d pre-order rank
//    from rankpos[pos]; the FIFO holds lids and an entry's coordinates / cell key come from the component-contiguous
//    copies mpts[t], mkey[t] (prefetched into L1 when the entry is pushed). A BFS frontier re-reads the same cells round
//    after round: those reads now hit L1 instead of making an L2 round trip each (the first generation re-read
//    1.07 GB from DRAM per 154-frame launch, 8x its compulsory bytes),
//  * labels[k] = label (seed_of) is written once per point, when it is removed: every point a BFS touches is removed
//    before that BFS ends (clustering.cpp:85-105), so the last write of the reference is the only one that matters,
//  * a candidate's distance is tested BEFORE its state: only members of the component can pass (r-components are closed
//    under the radius test), so the state lookup never leaves the component's bitmap.
//
// Only the per-cell live counters (tlive, fire-and-forget atomics, read once per round in the lookup phase) stay
// mutable in global memory.
#pragma once

#include "replay_cta.cuh"

namespace lb
{

constexpr uint32_t kPoolCap = 4096u; // candidates of one round (= push + postponed-removal slots) kept in shared memory

struct __align__(16) Cta2Smem
{
    uint32_t ring[kRing];               // lids of the most recent FIFO entries
    unsigned long long pool[kPoolCap];  // entry k owns [seg[k], seg[k+1]): its pushes (rank << 32 | lid) from the front,
                                        // its postponed removals (lid) from the back (a candidate is one or the other)
    float4 ent[kCtaW];                  // entry coordinates (w = bits(lid))
    unsigned long long ent_key[kCtaW];  // cell key of the entry
    uint32_t ent_widx[kCtaW];           // window index of the entry
    uint32_t cstart[kCtaW][27], cincl[kCtaW][27], cslot[kCtaW][27]; // the 27 neighbour cells of entry k: first pos,
                                        // inclusive candidate prefix, hash slot (live counter)
    uint32_t tk[kCtaW], np[kCtaW], nd[kCtaW]; // candidates / pushes / postponed removals of entry k
    uint32_t seg[kCtaW + 1u];                 // exclusive prefix of tk[] (every warp writes the same values)
    uint32_t wcnt[8];
    uint32_t claim, found;
};

// component state bitmap: 2 bits per member (kStRemoved | kStQueued), 16 members per word
LB_D uint32_t st_get(const volatile uint32_t *st, uint32_t lid)
{
    return (st[lid >> 4] >> ((lid & 15u) << 1)) & 3u;
}
LB_D void st_or(uint32_t *st, uint32_t lid, uint32_t bits)
{
    atomicOr(&st[lid >> 4], bits << ((lid & 15u) << 1));
}
LB_D void prefetch_l1(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(__cvta_generic_to_global(p)));
}

// Job lists of the CTA paths. Buckets 0..3: components of at most `normal_max` members, by size (longest first);
// bucket 4: "huge" components (bitmap of the 1-CTA-per-SM launch); bucket 5: beyond that, replayed by the
// first-generation kernel whose state lives in global memory. Counter layout behind m_cursor(): see api.cu.
constexpr uint32_t kBigListBuckets = 6u;
constexpr uint32_t kReplay2NormalWords = 2048u;  // 8 KB of state: components of up to 32 768 members, 3 CTAs per SM
constexpr uint32_t kReplay2HugeWords = 40960u;   // 160 KB of state: up to 655 360 members, 1 CTA per SM

__global__ void __launch_bounds__(256)
replay_biglist2_kernel(BatchView bv, const uint32_t *__restrict__ member_root, const uint32_t *__restrict__ comp_size,
                       uint32_t cta_min_members, uint32_t normal_max, uint32_t huge_max, uint2 *__restrict__ biglist,
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
                if (size > huge_max)
                    biglist[5u * bucket_capacity + atomicAdd(legacy_count, 1u)] = make_uint2(f, t);
                else if (size > normal_max)
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

// ipts[pos] = {x, y, z, bits(member slot t)}, rankpos[pos] = k-d pre-order rank, mpts[t] = {x, y, z, bits(pos)},
// mkey[t] = cell key: the immutable working set of the second-generation CTA replay. One thread per member slot.
__global__ void __launch_bounds__(256)
replay_init2_kernel(const float4 *__restrict__ cpts, BatchView bv, const uint32_t *__restrict__ rank_of_point,
                    const uint32_t *__restrict__ member_idx, const uint32_t *__restrict__ pos_of,
                    const unsigned long long *__restrict__ pkey, float4 *__restrict__ ipts, uint32_t *__restrict__ rankpos,
                    float4 *__restrict__ mpts, unsigned long long *__restrict__ mkey)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x)
    {
        const uint32_t idx = member_idx[off + t];
        const uint32_t pos = pos_of[off + idx];
        const float4 p = cpts[off + pos];
        ipts[off + pos] = make_float4(p.x, p.y, p.z, __uint_as_float(t));
        rankpos[off + pos] = rank_of_point[off + idx];
        mpts[off + t] = make_float4(p.x, p.y, p.z, __uint_as_float(pos));
        mkey[off + t] = pkey[off + pos];
    }
}

// MINB = CTAs per SM the register allocation is capped for. Dynamic shared memory: Cta2Smem followed by
// `state_words` words of component state; the job lists only hold components of at most 16 * state_words members.
template <int MINB>
__global__ void __launch_bounds__(kCtaThreads, MINB)
replay_cta2_kernel(const float4 *__restrict__ ipts_all, const uint32_t *__restrict__ rankpos_all,
                   const float4 *__restrict__ mpts_all, const unsigned long long *__restrict__ mkey_all, BatchView bv,
                   TableView tv, const uint4 *__restrict__ cells, CluParams prm, const uint32_t *__restrict__ member_root,
                   const uint32_t *__restrict__ member_idx, const uint32_t *__restrict__ comp_size,
                   uint32_t *__restrict__ tlive_all, uint32_t *__restrict__ seed_of, uint32_t *__restrict__ queue,
                   unsigned long long *__restrict__ push_spill, uint8_t *__restrict__ seed_valid,
                   const uint2 *__restrict__ biglist, uint32_t bucket_capacity, const uint32_t *__restrict__ big_count,
                   uint32_t n_buckets, uint32_t *__restrict__ cursor, uint32_t state_words,
                   uint32_t *__restrict__ job_stats /* optional: 8 words per job */)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cta2Smem &sm = *reinterpret_cast<Cta2Smem *>(smem_raw);
    uint32_t *st = reinterpret_cast<uint32_t *>(smem_raw + sizeof(Cta2Smem));
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const uint32_t lt = lanemask_lt();
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
        const uint32_t mask = table_mask(m, tv.tcap[f]);
        const uint4 *tab = cells + tv.toff[f];
        uint32_t *tlive = tlive_all + tv.toff[f];
        const float4 *ip = ipts_all + off;
        const uint32_t *rkp = rankpos_all + off;
        const float4 *mp = mpts_all + off + t_start;              // the component's members, by lid
        const unsigned long long *mk = mkey_all + off + t_start;
        uint32_t *so = seed_of + off;
        uint32_t *qu = queue + off + t_start; // the component's FIFO (lids)
        unsigned long long *spill = push_spill + off + t_start;
        const uint32_t *midx = member_idx + off + t_start;
        const uint32_t root = member_root[off + t_start];
        const uint32_t n_mem = comp_size[off + root];
        if (((n_mem + 15u) >> 4) > state_words)
            continue; // never listed (replay_biglist2_kernel routes by size); the labels would stay UNDEFINED

        for (uint32_t i = tid; i < ((n_mem + 15u) >> 4); i += kCtaThreads)
            st[i] = 0u;
        __syncthreads();

        const long long job_t0 = clock64();
        uint32_t st_rounds = 0u, st_direct = 0u, st_taken = 0u, st_seeds = 0u, st_cands = 0u;
        long long tA = 0, tB = 0, tC = 0, tmark = 0;
        uint32_t u = 0u; // next member (lid) to examine as a seed candidate (ascending index, clustering.cpp:70-75)
        while (true)
        {
            // ---- next seed: first member at or after u that is not removed
            uint32_t seed_l = 0xFFFFFFFFu;
            while (u < n_mem)
            {
                const uint32_t uu = u + tid;
                const bool cand = uu < n_mem && (st_get(st, uu) & kStRemoved) == 0u;
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
            ++st_seeds;
            u = seed_l + 1u;
            const uint32_t seed_idx = midx[seed_l];

            uint32_t head = 0u, tail = 1u, touched = 0u; // touched: this thread's share
            if (tid == 0)
            {
                qu[0] = seed_l;
                sm.ring[0] = seed_l;
                st_or(st, seed_l, kStQueued);
            }
            __syncthreads();

            while (head < tail) // clustering.cpp:80-111
            {
                // ---- A: window of the next 256 FIFO entries, the first kCtaW live ones are taken
                tmark = clock64();
                const uint32_t e = head + tid;
                uint32_t elid = 0u;
                bool alive = false;
                if (e < tail)
                {
                    elid = (tail - e <= kRing) ? sm.ring[e & (kRing - 1u)] : __ldcg(&qu[e]);
                    alive = (st_get(st, elid) & kStRemoved) == 0u;
                }
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
                    float4 pe = __ldg(&mp[elid]);
                    pe.w = __uint_as_float(elid);
                    sm.ent[arank] = pe;
                    sm.ent_key[arank] = __ldg(&mk[elid]);
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
                {
                    uint32_t T = 0u;
                    if (mine)
                    {
                        uint32_t start = 0u, count = 0u, slot = 0u;
                        if (lane < 27u)
                        {
                            // neighbour key = own key + (dx, dy, dz) in the packed 21-bit fields (biased, no borrow)
                            const long long dk = static_cast<long long>(static_cast<int>(lane % 3u) - 1) +
                                                 (static_cast<long long>(static_cast<int>((lane / 3u) % 3u) - 1) << 21) +
                                                 (static_cast<long long>(static_cast<int>(lane / 9u) - 1) << 42);
                            cell_lookup_alive(tab, tlive, mask, sm.ent_key[warp] + static_cast<unsigned long long>(dk), &start,
                                              &count, &slot);
                        }
                        const uint32_t incl = warp_inclusive_scan(count);
                        T = __shfl_sync(kFullMask, incl, 26);
                        if (lane < 27u)
                        {
                            sm.cstart[warp][lane] = start;
                            sm.cincl[warp][lane] = incl;
                            sm.cslot[warp][lane] = slot;
                        }
                        st_cands += T;
                    }
                    if (lane == 0)
                    {
                        sm.tk[warp] = T;
                        sm.np[warp] = 0u;
                        sm.nd[warp] = 0u;
                    }
                }
                __syncthreads();
                { const long long t = clock64(); tB += t - tmark; tmark = t; }
                // the round expands the longest prefix of the taken entries whose candidates fit the pool together; a
                // dense FIRST entry that does not fit is expanded alone and spills its pushes to global memory
                uint32_t n_use = 0u, seg[kCtaW + 1u];
                seg[0] = 0u;
#pragma unroll
                for (uint32_t v = 0; v < kCtaW; ++v)
                {
                    const uint32_t c = v < n_take ? sm.tk[v] : 0u;
                    seg[v + 1u] = seg[v] + c;
                    if (v < n_take && n_use == v && seg[v + 1u] <= kPoolCap)
                        n_use = v + 1u;
                }
                const bool spill_mode = n_use == 0u; // then entry 0 alone: tk[0] > kPoolCap, no earlier entry, nothing postponed
                if (spill_mode)
                    n_use = 1u;
                if (lane == 0)
                {
#pragma unroll
                    for (uint32_t v = 0; v <= kCtaW; ++v)
                        sm.seg[v] = seg[v];
                }
                __syncwarp();
                const uint32_t t_total = sm.seg[n_use];

                // ---- C: every candidate of the round is treated like the loop body of clustering.cpp:94-109 by ONE thread:
                // the (entry, candidate) pairs of all expanded entries are dealt over the whole CTA, so a round makes one
                // trip to the point records however its candidates are spread over the entries
                for (uint32_t base = 0; base < t_total; base += kCtaThreads * kCtaUnroll)
                {
                    uint32_t pos2[kCtaUnroll], slot2[kCtaUnroll], rank2[kCtaUnroll], ent2[kCtaUnroll], seg2[kCtaUnroll];
                    float4 cand2[kCtaUnroll];
                    bool valid2[kCtaUnroll];
#pragma unroll
                    for (int h = 0; h < kCtaUnroll; ++h)
                    {
                        const uint32_t g = base + kCtaThreads * h + tid;
                        valid2[h] = g < t_total;
                        cand2[h] = make_float4(0.f, 0.f, 0.f, 0.f);
                        pos2[h] = slot2[h] = rank2[h] = ent2[h] = seg2[h] = 0u;
                        if (valid2[h])
                        {
                            uint32_t k = 0u; // entry owning candidate g
#pragma unroll
                            for (uint32_t v = 1; v < kCtaW; ++v)
                                k += (v < n_use && g >= sm.seg[v]) ? 1u : 0u;
                            const uint32_t segk = sm.seg[k];
                            const uint32_t q = g - segk;
                            seg2[h] = segk;
                            const uint32_t *ci = sm.cincl[k];
                            uint32_t lo = 0u, hi = 26u;
#pragma unroll
                            for (int it = 0; it < 5; ++it) // first cell whose inclusive prefix exceeds q
                            {
                                const uint32_t mid = (lo + hi) >> 1;
                                if (ci[mid] > q)
                                    hi = mid;
                                else
                                    lo = mid + 1u;
                            }
                            ent2[h] = k;
                            pos2[h] = sm.cstart[k][lo] + (q - (lo ? ci[lo - 1u] : 0u));
                            slot2[h] = sm.cslot[k][lo];
                            cand2[h] = __ldg(&ip[pos2[h]]);
                            rank2[h] = __ldg(&rkp[pos2[h]]);
                        }
                    }
#pragma unroll
                    for (int h = 0; h < kCtaUnroll; ++h)
                    {
                        if (!valid2[h])
                            continue;
                        const float4 cand = cand2[h];
                        const uint32_t k = ent2[h];
                        const float4 pj = sm.ent[k];
                        // KDTree::dist_sqr(target, node) (kdtree.hpp:145-163), inclusive test (kdtree.hpp:314)
                        const float d2 = dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z);
                        const uint32_t lidc = __float_as_uint(cand.w) - t_start;
                        if (!(d2 <= prm.distance_squared) || lidc >= n_mem)
                            continue;
                        const uint32_t sw = st_get(st, lidc);
                        if (sw & kStRemoved) // removed points are skipped (clustering.cpp:94-97)
                            continue;
                        // what the entries expanded earlier in this round did to the candidate
                        bool removed_before = false, queued_before = (sw & kStQueued) != 0u;
                        bool shared = false; // an earlier entry of the round reaches the candidate too
                        for (uint32_t em = applied & ((1u << k) - 1u); em; em &= em - 1u)
                        {
                            const float4 po = sm.ent[__ffs(em) - 1];
                            const float dj = dist_sqr_ref(po.x, po.y, po.z, cand.x, cand.y, cand.z);
                            removed_before |= dj <= prm.inner_threshold;
                            shared |= dj <= prm.distance_squared;
                        }
                        queued_before |= shared;
                        if (removed_before)
                            continue;
                        ++touched; // indices_.push_back (with multiplicity)
                        if (d2 <= prm.inner_threshold)
                        {
                            // clustering.cpp:99,102-105: the point leaves the cloud with this seed's label. An EARLIER
                            // entry that reaches this candidate must still see it alive (it touches it first in the
                            // reference's order), whichever thread gets here first: the state write of such a removal waits
                            // for the CTA barrier; the label and the live counter are not read inside a round.
                            so[pos2[h]] = seed_idx;
                            atomicSub(&tlive[slot2[h]], 1u);
                            if (shared)
                                sm.pool[seg2[h] + sm.tk[k] - 1u - atomicAdd(&sm.nd[k], 1u)] = lidc;
                            else
                                st_or(st, lidc, kStRemoved);
                        }
                        else if (!queued_before)
                        {
                            st_or(st, lidc, kStQueued); // clustering.cpp:106-109 (first push only)
                            const uint32_t idx = atomicAdd(&sm.np[k], 1u);
                            const unsigned long long key =
                                (static_cast<unsigned long long>(rank2[h]) << 32) | static_cast<unsigned long long>(lidc);
                            if (idx < kPoolCap)
                                sm.pool[seg2[h] + idx] = key; // (the segment starts at 0 in spill mode)
                            else
                                spill[tail + idx] = key;
                        }
                    }
                }
                __syncthreads();
                { const long long t = clock64(); tC += t - tmark; tmark = t; }
                // ---- D: all state reads of the round are done; the postponed removals are written
                if (warp < n_use)
                {
                    const uint32_t seg_end = sm.seg[warp + 1u];
                    for (uint32_t i = lane; i < sm.nd[warp]; i += 32u)
                        st_or(st, static_cast<uint32_t>(sm.pool[seg_end - 1u - i]), kStRemoved);
                }
                // ---- F: the FIFO receives the pushes ordered by (entry, k-d pre-order rank)
                uint32_t np_total = 0u, seg_k = 0u;
                for (uint32_t k = 0; k < n_use; seg_k += sm.tk[k], ++k)
                {
                    const uint32_t npk = sm.np[k];
                    if (npk == 0u)
                        continue;
                    unsigned long long *pbuf = sm.pool + seg_k;
                    if (npk > kPoolCap)
                    {
                        // rare (spill mode only): sort in global memory, the spill area holds entries kPoolCap.. already
                        for (uint32_t i = tid; i < kPoolCap; i += kCtaThreads)
                            spill[tail + i] = sm.pool[i];
                        pbuf = spill + tail;
                        __syncthreads();
                    }
                    if (npk <= 64u)
                    {
                        if (warp == (k & 7u)) // short lists: one warp ranks every key by counting the smaller ones
                            for (uint32_t e2 = lane; e2 < npk; e2 += 32u)
                            {
                                const unsigned long long key = pbuf[e2];
                                uint32_t dest = 0u;
                                for (uint32_t x = 0; x < npk; ++x)
                                    dest += pbuf[x] < key ? 1u : 0u;
                                const uint32_t lid = static_cast<uint32_t>(key);
                                qu[tail + np_total + dest] = lid;
                                sm.ring[(tail + np_total + dest) & (kRing - 1u)] = lid;
                                prefetch_l1(&mp[lid]);
                                prefetch_l1(&mk[lid]);
                            }
                    }
                    else
                    {
                        cta_bitonic_sort(pbuf, npk); // (uniform branch: npk comes from shared memory)
                        for (uint32_t i = tid; i < npk; i += kCtaThreads)
                        {
                            const uint32_t lid = static_cast<uint32_t>(pbuf[i]);
                            qu[tail + np_total + i] = lid;
                            if (npk - i <= kRing)
                                sm.ring[(tail + np_total + i) & (kRing - 1u)] = lid;
                            if (i < 64u)
                            {
                                prefetch_l1(&mp[lid]);
                                prefetch_l1(&mk[lid]);
                            }
                        }
                    }
                    np_total += npk;
                }
                head += sm.ent_widx[n_use - 1u] + 1u;
                st_direct += spill_mode ? 1u : 0u;
                tail += np_total;
                __syncthreads();
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
            js[1] = n_mem;
            js[2] = static_cast<uint32_t>((clock64() - job_t0) >> 10);
            js[3] = st_rounds;
            js[4] = st_direct;
            js[5] = static_cast<uint32_t>(tA >> 10);
            js[6] = static_cast<uint32_t>(tB >> 10);
            js[7] = static_cast<uint32_t>(tC >> 10);
        }
    }
}

} // namespace lb
