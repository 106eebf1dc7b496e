// Window-synchronous replay of the reference BFS for large r-connected components (sm_100a), fifth generation.
//
// The order-dependent BFS of Clusterer::cluster (reference src/clustering.cpp:69-124) looks like a sequential chain —
// every pop depends on what the pops before it removed — but it has a closed form over a whole WINDOW of the FIFO:
//
//   Take the next n <= 256 FIFO entries (one per thread). An entry is expanded by the reference iff it is alive when it
//   is popped, i.e. iff it was alive at the start of the window and no EARLIER EXPANDED entry of the window holds it
//   within the inner radius (clustering.cpp:87-90, 102-105). The expanded entries are therefore the
//   lexicographically-first independent set of the "within the inner radius" conflict graph in queue order. It is
//   found without walking the queue: conflicts of all pairs from shared memory, then each warp settles its own 32
//   entries with ballots and the warps settle one after the other (<= 8 CTA steps, usually 1-3).
//   With the expanded set known, every (expanded entry j, candidate k) pair is independent work: against the state at
//   window start,
//        removed_before(j, k) = an expanded entry e < j holds k within its inner radius,
//        queued_before(j, k)  = k was queued at window start, or an expanded entry e < j reaches k at all
//                               (it removed or queued it),
//   both decided from geometry alone: the coordinates of the window sit in shared memory and only the expanded
//   entries within twice the radius of j can matter (a bit mask per entry, built with the conflicts). The pushes of
//   the window enter the FIFO ordered by (window position of the pusher, k-d pre-order rank), which is the
//   reference's order (clustering.cpp:92-109 with the radius_search order of kdtree.hpp:292-341).
//
// tools/gen_probe.cpp restates this on the CPU and checks it against the sequential loop (labels and the cluster
// sizes counted with multiplicity): 0 differences on the data frames and the synthetic shapes. The longest job of the
// 154-frame sequence (a 22 k-member component, 1 092 expansions) takes 138 windows instead of the 575 speculative
// rounds of the second generation; a window costs about as much as one of those rounds.
//
// Where the data lives: removed (committed), removed-in-this-window and queued are three bit planes of the component
// in SHARED memory, indexed by the member's local id; removals of a window go to the second plane and are folded into
// the first behind the barrier that ends the window through a list of the words they touched, so every thread of the
// window reads the state of the window start. Point records are immutable and read through the read-only path
// (ipts / rankpos by cell order, mpts / mcell by member); the 27 neighbour cells of an entry come from the packed
// neighbour table (cc_nbr_kernel) in one load per cell. (Per-cell live counters that skip the cells behind the frontier
// were tried: the extra dependent load per lookup costs more than the skipped candidates save, 6.4 against 6.0 ms.)
#pragma once

#include "replay_cta2.cuh"

namespace lb
{

constexpr uint32_t kGenW = 256u;      // window: one FIFO entry per thread
constexpr uint32_t kGenPool = 2048u;  // pushes of one window kept in shared memory
constexpr uint32_t kGenDirty = 2048u; // state words touched by the removals of one window
constexpr uint32_t kGenPosBits = 20u; // nb27 packs first pos | count << 20
constexpr uint32_t kGenCountCap = 4095u;
constexpr uint32_t kGenBatch = 32u; // expanded entries whose neighbour cells are looked up together
constexpr int kGenUnroll = 2;       // candidates per lane in flight: a chunk is 32 * kGenUnroll candidates of one entry
constexpr uint32_t kGenChunk = 32u * kGenUnroll;
static_assert(kGenW == 256u, "the window masks are 8 words");

struct __align__(16) GenSmem
{
    float4 ent[kGenW];                 // window entries: coordinates, w = bits(lid)
    unsigned long long pool[kGenPool]; // pushes: window position << 56 | rank << 32 | lid
    uint32_t ent_cell[kGenW];          // cell id of the entry
    uint32_t conf[8][kGenW];           // per entry: earlier alive entries within the inner radius (word w of the window)
    uint32_t near_in[8][kGenW];        // per entry: earlier entries within twice the radius (expanded ones only after the settle)
    uint32_t ring[kRing];              // lids of the most recent FIFO entries
    uint32_t cst[kGenBatch][27], cin[kGenBatch][27]; // neighbour cells of the batch's expanded entries: first pos, inclusive prefix
    float box_lo[8][4], box_hi[8][4];  // bounding box of the alive entries of every 32-entry word of the window
    uint32_t in_mask[2][8], out_mask[2][8]; // expanded / not expanded, one bit per window entry; settle step s reads
                                            // copy s & 1 and writes the other one (no read races a write)
    uint32_t wcnt[32];
    uint32_t n_push, n_dirty, claim, found;
    uint16_t dirty[kGenDirty];
    uint8_t in_list[kGenW]; // window positions of the expanded entries, ascending
};

// ipts[pos] = {x, y, z, bits(member slot t)}, rankpos[pos] = k-d pre-order rank, mpts[t] = {x, y, z, bits(pos)},
// mcell[t] = cell id. One thread per member slot.
__global__ void __launch_bounds__(256)
replay_init5_kernel(const float4 *__restrict__ cpts, BatchView bv, const uint32_t *__restrict__ rank_of_point,
                    const uint32_t *__restrict__ member_idx, const uint32_t *__restrict__ pos_of,
                    const uint32_t *__restrict__ cell_of, float4 *__restrict__ ipts, uint32_t *__restrict__ rankpos,
                    float4 *__restrict__ mpts, uint32_t *__restrict__ mcell)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x)
    {
        const uint32_t idx = member_idx[off + t];
        const uint32_t pos = pos_of[off + idx];
        const float4 p = cpts[off + pos];
        const uint32_t cid = cell_of[off + pos];
        ipts[off + pos] = make_float4(p.x, p.y, p.z, __uint_as_float(t));
        rankpos[off + pos] = rank_of_point[off + idx];
        mpts[off + t] = make_float4(p.x, p.y, p.z, __uint_as_float(pos));
        mcell[off + t] = cid;
    }
}

LB_D void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(__cvta_generic_to_global(p)));
}

LB_D bool plane_get(const volatile uint32_t *pl, uint32_t lid)
{
    return (pl[lid >> 5] >> (lid & 31u)) & 1u;
}

// One chunk of kGenChunk candidates of the expanded entry at window position p (number k in the batch), starting at
// candidate g0: every (entry, candidate) pair is treated like the loop body of clustering.cpp:94-109 by one lane.
// Returns this lane's number of touched points (indices_.push_back, with multiplicity).
// (Tried and dropped, both cost more instructions than they saved: a byte-mark + running-maximum mapping from candidate
// to cell instead of the binary search, and warp-aggregated counter updates instead of per-lane shared-memory atomics.)
LB_D uint32_t gen_scan_chunk(GenSmem &sm, uint32_t p, uint32_t k, uint32_t g0, uint32_t lane, const float4 *__restrict__ ip,
                             const uint32_t *__restrict__ rkp, const uint32_t *rem, uint32_t *rnew, uint32_t *qd,
                             uint32_t *__restrict__ so, unsigned long long *__restrict__ spill, const CluParams &prm,
                             uint32_t t_start, uint32_t n_mem, uint32_t seed_idx, uint32_t tail)
{
    const float4 pj = sm.ent[p];
    const uint32_t pw = p >> 5;
    const uint32_t cstart = lane < 27u ? sm.cst[k][lane] : 0u;
    const uint32_t incl = sm.cin[k][lane < 27u ? lane : 26u];
    const uint32_t excl = __shfl_up_sync(kFullMask, incl, 1);
    const uint32_t T = __shfl_sync(kFullMask, incl, 26);
    uint32_t pos2[kGenUnroll], rank2[kGenUnroll];
    float4 cand2[kGenUnroll];
    bool valid2[kGenUnroll];
#pragma unroll
    for (int h = 0; h < kGenUnroll; ++h)
    {
        const uint32_t g = g0 + 32u * h + lane;
        valid2[h] = g < T;
        uint32_t lo = 0u, hi = 26u;
#pragma unroll
        for (int it = 0; it < 5; ++it) // first cell whose inclusive prefix exceeds g
        {
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t v = __shfl_sync(kFullMask, incl, mid);
            if (v > g)
                hi = mid;
            else
                lo = mid + 1u;
        }
        lo = min(lo, 26u);
        const uint32_t cs = __shfl_sync(kFullMask, cstart, lo);
        const uint32_t ce = __shfl_sync(kFullMask, excl, lo); // (lane 0 holds its own inclusive value: unused)
        pos2[h] = cs + (g - (lo ? ce : 0u));
        cand2[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        rank2[h] = 0u;
        if (valid2[h])
        {
            cand2[h] = __ldg(&ip[pos2[h]]);
            rank2[h] = __ldg(&rkp[pos2[h]]);
        }
    }
    uint32_t touched = 0u;
#pragma unroll
    for (int h = 0; h < kGenUnroll; ++h)
    {
        if (!valid2[h])
            continue;
        const float4 cand = cand2[h];
        // KDTree::dist_sqr(target, node) (kdtree.hpp:145-163), inclusive test (kdtree.hpp:314)
        const float d2 = dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z);
        const uint32_t lidc = __float_as_uint(cand.w) - t_start;
        if (!(d2 <= prm.distance_squared) || lidc >= n_mem)
            continue;
        if (plane_get(rem, lidc)) // removed points are skipped (clustering.cpp:94-97)
            continue;
        // what the entries expanded earlier in this window did to the candidate
        bool removed_before = false, shared = false;
#pragma unroll 1
        for (uint32_t w = 0; w <= pw; ++w)
        {
            uint32_t em = sm.near_in[w][p];
            while (em)
            {
                const uint32_t b = static_cast<uint32_t>(__ffs(em) - 1);
                em &= em - 1u;
                const float4 po = sm.ent[w * 32u + b];
                const float dj = dist_sqr_ref(po.x, po.y, po.z, cand.x, cand.y, cand.z);
                removed_before |= dj <= prm.inner_threshold;
                shared |= dj <= prm.distance_squared;
            }
        }
        if (removed_before)
            continue;
        ++touched; // indices_.push_back (with multiplicity)
        if (d2 <= prm.inner_threshold)
        {
            // clustering.cpp:99,102-105: the point leaves the cloud with this seed's label. The bit goes to the plane of
            // this window: every pair of the window is judged against the state at its start.
            so[pos2[h]] = seed_idx;
            const uint32_t old = atomicOr(&rnew[lidc >> 5], 1u << (lidc & 31u));
            if (old == 0u)
            {
                const uint32_t di = atomicAdd(&sm.n_dirty, 1u);
                if (di < kGenDirty)
                    sm.dirty[di] = static_cast<uint16_t>(lidc >> 5);
            }
        }
        else if (!shared && !plane_get(qd, lidc))
        {
            // clustering.cpp:106-109 (first push only: a later copy of the entry is a no-op when popped)
            atomicOr(&qd[lidc >> 5], 1u << (lidc & 31u));
            const unsigned long long key = (static_cast<unsigned long long>(p) << 56) |
                                           (static_cast<unsigned long long>(rank2[h]) << 32) |
                                           static_cast<unsigned long long>(lidc);
            const uint32_t idx = atomicAdd(&sm.n_push, 1u);
            if (idx < kGenPool)
                sm.pool[idx] = key;
            else
                spill[tail + idx] = key;
        }
    }
    return touched;
}

// MINB = CTAs per SM the register allocation is capped for. Dynamic shared memory: GenSmem followed by three planes of
// `plane_words` words; the job lists only hold components of at most 32 * plane_words members.
// NT = threads per CTA: the first 256 own the window entries; tiles, lookups, candidate chunks and the sort are dealt over
// all NT / 32 warps, so the long jobs (whose windows are the chain that bounds the launch, and a single frame's latency)
// run with 1024 threads on an SM of their own while the many short ones share SMs with 256.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
replay_gen_kernel(const float4 *__restrict__ ipts_all, const uint32_t *__restrict__ rankpos_all,
                  const float4 *__restrict__ mpts_all, const uint32_t *__restrict__ mcell_all,
                  const uint32_t *__restrict__ nb27_all, const uint2 *__restrict__ cinfo_all, BatchView bv, CluParams prm,
                  const uint32_t *__restrict__ member_root, const uint32_t *__restrict__ member_idx,
                  const uint32_t *__restrict__ comp_size, uint32_t *__restrict__ seed_of, uint32_t *__restrict__ queue,
                  unsigned long long *__restrict__ push_spill, uint8_t *__restrict__ seed_valid,
                  const uint2 *__restrict__ biglist, uint32_t bucket_capacity, const uint32_t *__restrict__ big_count,
                  uint32_t n_buckets, uint32_t *__restrict__ cursor, uint32_t plane_words,
                  uint32_t *__restrict__ job_stats /* optional: 8 words per job */,
                  uint32_t stats_skip_buckets /* job lists in front of `big_count` that belong to another launch */)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GenSmem &sm = *reinterpret_cast<GenSmem *>(smem_raw);
    uint32_t *rem = reinterpret_cast<uint32_t *>(smem_raw + sizeof(GenSmem)); // removed before this window
    uint32_t *rnew = rem + plane_words;                                       // removed by this window
    uint32_t *qd = rnew + plane_words;                                        // queued (ever pushed)
    constexpr uint32_t kWarps = NT / 32;
    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31u;
    const uint32_t warp = tid >> 5;
    const bool owner = tid < kGenW; // this thread may own a window entry
    const uint32_t lt = lanemask_lt();
    const float near_sq = __fmul_rn(4.01f, prm.distance_squared); // superset of "within twice the radius"
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
        const uint32_t w_job = sm.claim;
        if (w_job >= n_big)
            break;
        uint32_t jb = 0u;
        while (w_job >= bucket_end[jb])
            ++jb;
        const uint2 job = biglist[jb * bucket_capacity + (w_job - (jb ? bucket_end[jb - 1u] : 0u))];
        const uint32_t f = job.x;
        const uint32_t t_start = job.y;
        const uint32_t off = bv.off[f];
        const float4 *ip = ipts_all + off;
        const uint32_t *rkp = rankpos_all + off;
        const float4 *mp = mpts_all + off + t_start; // the component's members, by lid
        const uint32_t *mc = mcell_all + off + t_start;
        const uint32_t *nbt = nb27_all + static_cast<size_t>(off) * 27u;
        const uint2 *ci = cinfo_all + off;
        uint32_t *so = seed_of + off;
        uint32_t *qu = queue + off + t_start; // the component's FIFO (lids)
        unsigned long long *spill = push_spill + off + t_start;
        const uint32_t *midx = member_idx + off + t_start;
        const uint32_t root = member_root[off + t_start];
        const uint32_t n_mem = comp_size[off + root];
        const uint32_t n_words = (n_mem + 31u) >> 5;
        if (n_words > plane_words)
            continue; // never listed (the job lists are routed by size); the labels would stay UNDEFINED

        for (uint32_t i = tid; i < n_words; i += NT)
        {
            rem[i] = 0u;
            rnew[i] = 0u;
            qd[i] = 0u;
        }
        if (tid == 0)
        {
            sm.n_push = 0u;
            sm.n_dirty = 0u;
        }
        __syncthreads();

        const long long job_t0 = clock64();
        uint32_t st_windows = 0u, st_in = 0u, st_entries = 0u, st_steps = 0u;
        long long t_load = 0, t_mis = 0, t_cand = 0, t_sort = 0, t_tiles = 0, t_settle = 0, t_look = 0, t_commit = 0;
        uint32_t u = 0u; // next member (lid) to examine as a seed candidate (ascending index, clustering.cpp:70-75)
        while (true)
        {
            // ---- next seed: first member at or after u that is not removed
            uint32_t seed_l = 0xFFFFFFFFu;
            for (uint32_t wbase = u >> 5; wbase < n_words; wbase += NT)
            {
                const uint32_t wi = wbase + tid;
                uint32_t cand = 0xFFFFFFFFu;
                if (wi < n_words)
                {
                    uint32_t freeb = ~rem[wi];
                    if (wi == (u >> 5))
                        freeb &= ~((1u << (u & 31u)) - 1u);
                    if (wi == n_words - 1u && (n_mem & 31u))
                        freeb &= (1u << (n_mem & 31u)) - 1u;
                    if (freeb)
                        cand = (wi << 5) + static_cast<uint32_t>(__ffs(freeb) - 1);
                }
                cand = warp_reduce_min(cand);
                if (tid == 0)
                    sm.found = 0xFFFFFFFFu;
                __syncthreads();
                if (lane == 0 && cand != 0xFFFFFFFFu)
                    atomicMin(&sm.found, cand);
                __syncthreads();
                seed_l = sm.found;
                __syncthreads();
                if (seed_l != 0xFFFFFFFFu)
                    break;
            }
            if (seed_l == 0xFFFFFFFFu)
                break; // component done
            u = seed_l + 1u;
            const uint32_t seed_idx = midx[seed_l];

            uint32_t head = 0u, tail = 1u, touched = 0u; // touched: this thread's share
            if (tid == 0)
            {
                qu[0] = seed_l;
                sm.ring[0] = seed_l;
                qd[seed_l >> 5] |= 1u << (seed_l & 31u);
            }
            __syncthreads();

            while (head < tail) // clustering.cpp:80-111
            {
                // ---- A: the window, one entry per thread
                const long long tw0 = clock64();
                const uint32_t n = min(kGenW, tail - head);
                ++st_windows;
                st_entries += n;
                bool alive = false;
                if (tid < n)
                {
                    const uint32_t e = head + tid;
                    const uint32_t lid = (tail - e <= kRing) ? sm.ring[e & (kRing - 1u)] : __ldcg(&qu[e]);
                    alive = !plane_get(rem, lid);
                    if (alive)
                    {
                        float4 pe = __ldg(&mp[lid]);
                        pe.w = __uint_as_float(lid);
                        sm.ent[tid] = pe;
                        const uint32_t cid = __ldg(&mc[lid]);
                        sm.ent_cell[tid] = cid;
                        prefetch_l2(&nbt[static_cast<size_t>(cid) * 27u]); // the row is wanted after the settle, if at all
                        prefetch_l2(&nbt[static_cast<size_t>(cid) * 27u + 26u]);
                    }
                }
                const uint32_t alive_w = __ballot_sync(kFullMask, alive);
                if (lane == 0 && owner)
                {
                    sm.in_mask[0][warp] = 0u;
                    sm.out_mask[0][warp] = ~alive_w;
                }
                if (warp * 32u < n && n > 32u) // bounding box of the word's alive entries: far-apart words skip their tile
                {
                    const float inf = __int_as_float(0x7f800000);
                    const float4 pe = alive ? sm.ent[tid] : make_float4(inf, inf, inf, 0.f);
                    float lx = pe.x, ly = pe.y, lz = pe.z;
                    float hx = alive ? pe.x : -inf, hy = alive ? pe.y : -inf, hz = alive ? pe.z : -inf;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1)
                    {
                        lx = fminf(lx, __shfl_xor_sync(kFullMask, lx, d));
                        ly = fminf(ly, __shfl_xor_sync(kFullMask, ly, d));
                        lz = fminf(lz, __shfl_xor_sync(kFullMask, lz, d));
                        hx = fmaxf(hx, __shfl_xor_sync(kFullMask, hx, d));
                        hy = fmaxf(hy, __shfl_xor_sync(kFullMask, hy, d));
                        hz = fmaxf(hz, __shfl_xor_sync(kFullMask, hz, d));
                    }
                    if (lane == 0)
                    {
                        sm.box_lo[warp][0] = lx;
                        sm.box_lo[warp][1] = ly;
                        sm.box_lo[warp][2] = lz;
                        sm.box_hi[warp][0] = hx;
                        sm.box_hi[warp][1] = hy;
                        sm.box_hi[warp][2] = hz;
                    }
                }
                __syncthreads();
                const long long tw1 = clock64();
                t_load += tw1 - tw0;

                // ---- conflicts (inner radius) of every entry with the EARLIER entries: 32 x 32 tiles of the pair matrix, dealt
                // over the warps; lane = the later entry. A tile whose two words lie further apart than the inner radius
                // (bounding boxes; a BFS frontier queues neighbours next to each other) is all zero. Dead entries hold stale
                // coordinates: their bits are masked with the alive words.
                const uint32_t nw = (n + 31u) >> 5;
                if (n > 1u)
                {
                    uint32_t wp = 0u, we = warp; // tile number `warp` of the lower triangle, row by row
                    while (we > wp)
                    {
                        we -= wp + 1u;
                        ++wp;
                    }
                    while (wp < nw)
                    {
                        bool far = false;
                        if (we != wp)
                        {
                            float gap2 = 0.f;
#pragma unroll
                            for (int a = 0; a < 3; ++a)
                            {
                                const float g = fmaxf(fmaxf(sm.box_lo[wp][a] - sm.box_hi[we][a], sm.box_lo[we][a] - sm.box_hi[wp][a]), 0.f);
                                gap2 += g * g;
                            }
                            far = !(gap2 <= 1.001f * prm.inner_threshold); // (an empty word has an infinite gap; NaN never prunes... it cannot occur)
                        }
                        uint32_t cw = 0u;
                        if (!far)
                        {
                            const float4 me = sm.ent[wp * 32u + lane];
                            const uint32_t cnt_e = min(32u, n - we * 32u);
#pragma unroll 4
                            for (uint32_t b = 0; b < cnt_e; ++b)
                            {
                                const float4 q = sm.ent[we * 32u + b];
                                const float d2 = dist_sqr_ref(q.x, q.y, q.z, me.x, me.y, me.z);
                                cw |= (d2 <= prm.inner_threshold ? 1u : 0u) << b;
                            }
                            uint32_t keep = ~sm.out_mask[0][we]; // (only dead entries are in out_mask before the settle starts)
                            if (we == wp)
                                keep &= lt;
                            cw &= keep;
                        }
                        sm.conf[we][wp * 32u + lane] = cw;
                        for (uint32_t adv = 0; adv < kWarps; ++adv) // next tile of this warp
                            if (++we > wp)
                            {
                                we = 0u;
                                ++wp;
                            }
                    }
                }
                __syncthreads();
                const long long tw2 = clock64();
                t_tiles += tw2 - tw1;

                // ---- which entries are expanded: the lexicographically-first independent set. A warp settles its own
                // entries with ballots; warp w is settled once the warps before it are, i.e. after at most w + 1 steps.
                bool unresolved = alive, is_in = false;
                uint32_t fin = 0u; // which copy of the masks the next settle step reads; the final one afterwards
                {
                    // conflicts with the entries of earlier warps, of the own warp, and the earlier words that matter to this
                    // warp at all (uniform): a BFS frontier queues neighbours next to each other, most words hold none
                    uint32_t conf[8], own_conf = 0u, wmask = 0u;
#pragma unroll
                    for (int w = 0; w < 8; ++w)
                    {
                        conf[w] = 0u;
                        if (owner && n > 1u && static_cast<uint32_t>(w) <= warp && static_cast<uint32_t>(w) < nw) // (uniform per warp)
                        {
                            const uint32_t v = alive ? sm.conf[w][tid] : 0u;
                            if (static_cast<uint32_t>(w) == warp)
                                own_conf = v;
                            else
                            {
                                conf[w] = v;
                                wmask |= __any_sync(kFullMask, v != 0u) ? 1u << w : 0u;
                            }
                        }
                    }
                    uint32_t own_in = 0u, own_out = ~alive_w, st_guard = 0u;
                    while (true)
                    {
                        const volatile uint32_t *vin = sm.in_mask[fin], *vout = sm.out_mask[fin];
                        fin ^= 1u;
                        ++st_steps;
                        // what the earlier warps have settled so far (a snapshot per step is enough: a warp is settled one
                        // step after the last warp it depends on)
                        uint32_t any_in_prev = 0u, pend_prev = 0u;
#pragma unroll
                        for (int w = 0; w < 8; ++w)
                            if ((wmask >> w) & 1u)
                            {
                                const uint32_t iw = vin[w], ow = vout[w];
                                any_in_prev |= conf[w] & iw;
                                pend_prev |= conf[w] & ~(iw | ow);
                            }
                        uint32_t moved;
                        do
                        {
                            bool new_in = false, new_out = false;
                            if (unresolved)
                            {
                                const uint32_t any_in = any_in_prev | (own_conf & own_in);
                                const uint32_t pending = pend_prev | (own_conf & ~(own_in | own_out));
                                new_out = any_in != 0u;
                                new_in = any_in == 0u && pending == 0u;
                            }
                            const uint32_t bi = __ballot_sync(kFullMask, new_in);
                            const uint32_t bo = __ballot_sync(kFullMask, new_out);
                            own_in |= bi;
                            own_out |= bo;
                            is_in = is_in || new_in;
                            unresolved = unresolved && !new_in && !new_out;
                            moved = bi | bo;
                        } while (moved);
                        if (lane == 0 && owner)
                        {
                            sm.in_mask[fin][warp] = own_in;
                            sm.out_mask[fin][warp] = own_out;
                        }
                        if (!__syncthreads_or(unresolved ? 1 : 0) || st_guard++ > 16u) // (settled after <= 8 steps by construction)
                            break;
                    }
                }
                // (the barrier above made every warp's final masks visible)
                t_settle += clock64() - tw2;
                uint32_t n_in = 0u;
                {
                    uint32_t before = 0u;
#pragma unroll
                    for (uint32_t w = 0; w < 8u; ++w)
                    {
                        const uint32_t c = __popc(sm.in_mask[fin][w]);
                        before += w < warp ? c : 0u;
                        n_in += c;
                    }
                    if (is_in)
                        sm.in_list[before + __popc(sm.in_mask[fin][warp] & lt)] = static_cast<uint8_t>(tid);
                }
                st_in += n_in;
                __syncthreads();
                // per expanded entry: the earlier expanded entries within twice the radius (the only ones that can have touched
                // one of its candidates). Thread t owns column in_list[t]; visible to the scans behind the lookup barrier.
                if (tid < n_in)
                {
                    const uint32_t p = sm.in_list[tid];
                    const float4 me = sm.ent[p];
                    uint32_t acc = 0u, cur_w = 0u;
                    for (uint32_t e = 0; e < tid; ++e)
                    {
                        const uint32_t pe = sm.in_list[e];
                        if ((pe >> 5) != cur_w)
                        {
                            sm.near_in[cur_w][p] = acc;
                            for (uint32_t w = cur_w + 1u; w < (pe >> 5); ++w)
                                sm.near_in[w][p] = 0u;
                            cur_w = pe >> 5;
                            acc = 0u;
                        }
                        const float4 q = sm.ent[pe];
                        acc |= (dist_sqr_ref(q.x, q.y, q.z, me.x, me.y, me.z) <= near_sq ? 1u : 0u) << (pe & 31u);
                    }
                    sm.near_in[cur_w][p] = acc;
                    for (uint32_t w = cur_w + 1u; w <= (p >> 5); ++w)
                        sm.near_in[w][p] = 0u;
                }
                const long long tc0 = clock64();
                t_mis += tc0 - tw1;

                // ---- candidates, in batches of kGenBatch expanded entries: the 27 neighbour cells of every entry of the batch are
                // looked up (warp per entry), then the candidates of the batch are dealt over the warps in chunks
                for (uint32_t b0 = 0; b0 < n_in; b0 += kGenBatch)
                {
                    const uint32_t nb = min(kGenBatch, n_in - b0);
                    for (uint32_t k = warp; k < nb; k += kWarps)
                    {
                        const uint32_t cid = sm.ent_cell[sm.in_list[b0 + k]];
                        uint32_t cstart = 0u, count = 0u;
                        if (lane < 27u)
                        {
                            const uint32_t v = __ldg(&nbt[static_cast<size_t>(cid) * 27u + lane]);
                            cstart = v & ((1u << kGenPosBits) - 1u);
                            count = v >> kGenPosBits;
                            if (count == kGenCountCap)
                                count = __ldg(&ci[cstart]).x;
                        }
                        const uint32_t incl = warp_inclusive_scan(count);
                        if (lane < 27u)
                        {
                            sm.cst[k][lane] = cstart;
                            sm.cin[k][lane] = incl;
                        }
                    }
                    __syncthreads();
                    const long long tl1 = clock64();
                    t_look += tl1 - (b0 == 0u ? tc0 : tl1);
                    {
                        // lane k: chunks of entry k of the batch; chunk c belongs to the first entry whose inclusive prefix exceeds c
                        const uint32_t my_chunks = lane < nb ? (sm.cin[lane][26] + kGenChunk - 1u) / kGenChunk : 0u;
                        const uint32_t cincl = warp_inclusive_scan(my_chunks);
                        const uint32_t n_chunks = __shfl_sync(kFullMask, cincl, 31);
                        for (uint32_t c = warp; c < n_chunks; c += kWarps)
                        {
                            const uint32_t k = static_cast<uint32_t>(__ffs(__ballot_sync(kFullMask, cincl > c)) - 1);
                            const uint32_t first = __shfl_sync(kFullMask, cincl - my_chunks, k);
                            touched += gen_scan_chunk(sm, sm.in_list[b0 + k], k, (c - first) * kGenChunk, lane, ip, rkp, rem, rnew, qd, so,
                                                      spill, prm, t_start, n_mem, seed_idx, tail);
                        }
                    }
                    if (b0 + kGenBatch < n_in)
                        __syncthreads(); // the next batch overwrites the cell tables
                }
                __syncthreads();
                const long long tc1 = clock64();
                t_cand += tc1 - tc0;

                // ---- the removals of the window are committed
                {
                    const uint32_t nd = sm.n_dirty;
                    if (nd <= kGenDirty)
                        for (uint32_t i = tid; i < nd; i += NT)
                        {
                            const uint32_t wd = sm.dirty[i];
                            const uint32_t bits = rnew[wd];
                            rem[wd] |= bits;
                            rnew[wd] = 0u;
                        }
                    else
                        for (uint32_t i = tid; i < n_words; i += NT)
                        {
                            const uint32_t bits = rnew[i];
                            if (bits == 0u)
                                continue;
                            rem[i] |= bits;
                            rnew[i] = 0u;
                        }
                }
                t_commit += clock64() - tc1;
                // ---- the FIFO receives the pushes ordered by (window position of the pusher, k-d pre-order rank)
                const uint32_t np = sm.n_push;
                if (tail + np > n_mem) // cannot happen (a member is pushed once): never write past the component's FIFO
                    break;
                if (np)
                {
                    if (np <= static_cast<uint32_t>(NT))
                    {
                        // short lists: every key is ranked by counting the smaller ones; up to 16 adjacent lanes share a key
                        uint32_t parts = 1u;
                        while (parts < 16u && 2u * parts * np <= static_cast<uint32_t>(NT))
                            parts <<= 1;
                        const uint32_t i = tid / parts, part = tid % parts;
                        const uint32_t len = (np + parts - 1u) / parts;
                        const uint32_t xb = min(np, part * len), xe = min(np, xb + len);
                        const unsigned long long key = sm.pool[min(i, np - 1u)];
                        uint32_t dest = 0u;
#pragma unroll 8
                        for (uint32_t x = xb; x < xe; ++x)
                            dest += sm.pool[x] < key ? 1u : 0u;
                        for (uint32_t o = 1u; o < parts; o <<= 1) // (uniform) the shares of the adjacent lanes of a key
                            dest += __shfl_xor_sync(kFullMask, dest, o);
                        if (i < np && part == 0u)
                        {
                            const uint32_t lid = static_cast<uint32_t>(key);
                            qu[tail + dest] = lid;
                            sm.ring[(tail + dest) & (kRing - 1u)] = lid;
                            prefetch_l1(&mp[lid]);
                            prefetch_l1(&mc[lid]);
                        }
                    }
                    else
                    {
                        volatile unsigned long long *pbuf = sm.pool;
                        if (np > kGenPool)
                        {
                            // rare: sort in global memory, the spill area holds the keys from kGenPool on already
                            for (uint32_t i = tid; i < kGenPool; i += NT)
                                spill[tail + i] = sm.pool[i];
                            pbuf = spill + tail;
                            __syncthreads();
                        }
                        cta_bitonic_sort<NT>(pbuf, np); // (uniform branch: np comes from shared memory)
                        for (uint32_t i = tid; i < np; i += NT)
                        {
                            const uint32_t lid = static_cast<uint32_t>(pbuf[i]);
                            qu[tail + i] = lid;
                            if (np - i <= kRing)
                                sm.ring[(tail + i) & (kRing - 1u)] = lid;
                            if (i < kGenW)
                            {
                                prefetch_l1(&mp[lid]);
                                prefetch_l1(&mc[lid]);
                            }
                        }
                    }
                }
                head += n;
                tail += np;
                __syncthreads();
                t_sort += clock64() - tc1;
                if (tid == 0)
                {
                    sm.n_push = 0u;
                    sm.n_dirty = 0u;
                        }
                // (the next reader / writer of these counters sits behind the barrier that follows the window load)
            }
            // ---- seed finished: cluster size test with multiplicity (clustering.cpp:113-123)
            touched = warp_reduce_add(touched);
            if (lane == 0)
                sm.wcnt[warp] = touched;
            __syncthreads();
            if (tid == 0)
            {
                uint32_t tsum = 0u;
                for (uint32_t v = 0; v < kWarps; ++v)
                    tsum += sm.wcnt[v];
                seed_valid[off + seed_idx] = (tsum < prm.min_cluster_size || tsum > prm.max_cluster_size) ? 0u : 1u;
            }
            __syncthreads();
        }
        if (job_stats && tid == 0)
        {
            uint32_t skip = 0u;
            for (uint32_t b = 1; b <= stats_skip_buckets; ++b)
                skip += *(big_count - b);
            uint32_t *js = job_stats + 8u * (skip + w_job);
            js[0] = f;
            js[1] = n_mem;
            js[2] = static_cast<uint32_t>((clock64() - job_t0) >> 10);
            js[3] = st_windows | (st_in << 16);
            js[4] = (static_cast<uint32_t>(t_tiles >> 10) & 0xFFFFu) | (static_cast<uint32_t>(t_settle >> 10) << 16);
            js[5] = (static_cast<uint32_t>(t_look >> 10) & 0xFFFFu) | (static_cast<uint32_t>(t_commit >> 10) << 16);
            js[6] = (static_cast<uint32_t>(t_load >> 10) & 0xFFFFu) | (static_cast<uint32_t>(t_mis >> 10) << 16);
            js[7] = (static_cast<uint32_t>(t_cand >> 10) & 0xFFFFu) | (static_cast<uint32_t>(t_sort >> 10) << 16);
        }
    }
}

} // namespace lb
