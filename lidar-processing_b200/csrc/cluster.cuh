// Fast Euclidean Clustering kernels (sm_100a). Drop-in for lidar_processing::Clusterer::cluster
// (reference src/clustering.cpp:47-125) with the k-d tree radius search (src/kdtree.hpp:292-341)
// replaced by a voxel-hash grid whose cell edge is the cluster tolerance (27-cell neighbour scan).
//
// The reference result is NOT the connected components of the radius graph: it is an order-dependent
// BFS ("remove everything within (1-q)·r of an expanded point, enqueue the annulus") whose outcome
// depends on the order in which radius_search reports neighbours, i.e. the k-d tree pre-order. The
// device therefore
//   1. builds the voxel hash (insert -> scan -> fill),
//   2. finds the r-connected components with a lock-free union-find (atomicCAS hooking of the
//      larger root under the smaller + path halving): BFS runs never cross component borders, so
//      components are independent replay units,
//   3. replays the reference BFS inside every component with one warp: grid neighbours are tested
//      with the reference's exact float d², the points that enter the FIFO are ordered by their
//      k-d pre-order rank (kd_build.cuh), the FIFO is de-duplicated (a duplicate entry is a no-op
//      when popped in the reference, because the first pop always ends with the point removed),
//   4. compacts labels: label k = k-th valid seed in ascending seed index (clustering.cpp:113-123).
#pragma once

#include "common.cuh"

namespace lb
{

constexpr uint64_t kCellEmpty = 0xFFFFFFFFFFFFFFFFull;
constexpr uint32_t kStRemoved = 1u;
constexpr uint32_t kStQueued = 2u;
constexpr uint32_t kSeedUnset = 0xFFFFFFFFu;
constexpr int kCellBias = 1 << 20;
constexpr int32_t kLabelUndefined = static_cast<int32_t>(0x80000000u); // Clusterer::UNDEFINED (clustering.hpp:53)
constexpr int32_t kLabelInvalid = -1;                                   // Clusterer::INVALID   (clustering.hpp:54)

struct CluParams
{
    float distance_squared; // ClusteringConfiguration::distance_squared
    float inner_threshold;  // largest float t with (double)t <= (1-q)^2 * distance_squared (clustering.cpp:66-67)
    uint32_t min_cluster_size;
    uint32_t max_cluster_size;
    double inv_cell; // 1 / (sqrt(distance_squared) * 1.001)
    uint32_t cta_min_members; // components of at least this many members are replayed by a whole CTA (replay_cta.cuh)
};

// Per-frame placement of the hash table inside the batch-wide table arrays.
struct TableView
{
    const uint32_t *toff; // [F] first slot
    const uint32_t *tcap; // [F] reserved slots (power of two)
};

LB_D uint32_t table_mask(uint32_t m, uint32_t reserved)
{
    // effective capacity: smallest power of two >= 2*m (at least 64), never above the reservation
    uint32_t cap = 64u;
    while (cap < 2u * m && cap < reserved)
        cap <<= 1;
    return cap - 1u;
}

LB_D uint32_t hash_cell(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return static_cast<uint32_t>(k);
}

LB_D bool cell_coords(const float4 &p, double inv_cell, int *cx, int *cy, int *cz)
{
    const double fx = floor(static_cast<double>(p.x) * inv_cell);
    const double fy = floor(static_cast<double>(p.y) * inv_cell);
    const double fz = floor(static_cast<double>(p.z) * inv_cell);
    const double lim = static_cast<double>(kCellBias - 4);
    const bool ok = fabs(fx) < lim && fabs(fy) < lim && fabs(fz) < lim; // false for NaN/Inf too
    *cx = ok ? static_cast<int>(fx) + kCellBias : 0;
    *cy = ok ? static_cast<int>(fy) + kCellBias : 0;
    *cz = ok ? static_cast<int>(fz) + kCellBias : 0;
    return ok;
}

LB_D uint64_t cell_key(int cx, int cy, int cz)
{
    return static_cast<uint64_t>(static_cast<uint32_t>(cx)) | (static_cast<uint64_t>(static_cast<uint32_t>(cy)) << 21) |
           (static_cast<uint64_t>(static_cast<uint32_t>(cz)) << 42);
}

// cells[slot] = {key.lo, key.hi, start, count}
LB_D void cell_lookup(const uint4 *__restrict__ cells, uint32_t mask, uint64_t key, uint32_t *start, uint32_t *count)
{
    uint32_t slot = hash_cell(key) & mask;
    const uint32_t klo = static_cast<uint32_t>(key), khi = static_cast<uint32_t>(key >> 32);
    while (true)
    {
        const uint4 c = __ldg(&cells[slot]);
        if (c.x == klo && c.y == khi)
        {
            *start = c.z;
            *count = c.w;
            return;
        }
        if (c.x == 0xFFFFFFFFu && c.y == 0xFFFFFFFFu)
        {
            *start = 0u;
            *count = 0u;
            return;
        }
        slot = (slot + 1u) & mask;
    }
}

// like cell_lookup, also returning the slot (0xFFFFFFFF when the cell does not exist)
LB_D uint32_t cell_lookup_slot(const uint4 *__restrict__ cells, uint32_t mask, uint64_t key, uint32_t *start,
                               uint32_t *count)
{
    uint32_t slot = hash_cell(key) & mask;
    const uint32_t klo = static_cast<uint32_t>(key), khi = static_cast<uint32_t>(key >> 32);
    while (true)
    {
        const uint4 c = __ldg(&cells[slot]);
        if (c.x == klo && c.y == khi)
        {
            *start = c.z;
            *count = c.w;
            return slot;
        }
        if (c.x == 0xFFFFFFFFu && c.y == 0xFFFFFFFFu)
        {
            *start = 0u;
            *count = 0u;
            return 0xFFFFFFFFu;
        }
        slot = (slot + 1u) & mask;
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grid_clear_kernel(BatchView bv, TableView tv, unsigned long long *__restrict__ tkeys, uint32_t *__restrict__ tcount)
{
    const uint32_t f = blockIdx.y;
    const uint32_t cap = table_mask(bv.cnt[f], tv.tcap[f]) + 1u;
    const uint32_t toff = tv.toff[f];
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += gridDim.x * blockDim.x)
    {
        tkeys[toff + s] = kCellEmpty;
        tcount[toff + s] = 0u;
    }
}

__global__ void __launch_bounds__(256)
grid_insert_kernel(const float4 *__restrict__ pts, BatchView bv, TableView tv, CluParams prm,
                   unsigned long long *__restrict__ tkeys, uint32_t *__restrict__ tcount, uint32_t *__restrict__ slot_of,
                   uint32_t *__restrict__ err)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t mask = table_mask(m, tv.tcap[f]);
    const uint32_t toff = tv.toff[f];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    {
        const float4 p = pts[off + i];
        int cx, cy, cz;
        if (!cell_coords(p, prm.inv_cell, &cx, &cy, &cz))
            atomicOr(err, 1u); // non-finite or out-of-range coordinate
        const unsigned long long key = cell_key(cx, cy, cz);
        uint32_t slot = hash_cell(key) & mask;
        while (true)
        {
            const unsigned long long prev = atomicCAS(&tkeys[toff + slot], kCellEmpty, key);
            if (prev == kCellEmpty || prev == key)
                break;
            slot = (slot + 1u) & mask;
        }
        atomicAdd(&tcount[toff + slot], 1u);
        slot_of[off + i] = slot;
    }
}

// Exclusive scan of the per-slot counts in tiles of 4096 slots (grid = (tiles, frames)): pass 1 sums
// every tile, pass 2 adds the sums of the preceding tiles, emits the packed cell records and resets
// the counts (they become the fill cursors).
constexpr int kScanPer = 4;
constexpr uint32_t kScanTile = 1024u * kScanPer;

__global__ void __launch_bounds__(1024)
grid_scan_count_kernel(BatchView bv, TableView tv, const uint32_t *__restrict__ tcount, uint32_t max_tiles,
                       unsigned long long *__restrict__ tile_counts)
{
    __shared__ unsigned long long ws64[33];
    const uint32_t f = blockIdx.y;
    const uint32_t cap = table_mask(bv.cnt[f], tv.tcap[f]) + 1u;
    const uint32_t toff = tv.toff[f];
    const uint32_t first = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    unsigned long long sum = 0ull;
#pragma unroll
    for (int k = 0; k < kScanPer; ++k)
        sum += (first + k < cap) ? tcount[toff + first + k] : 0u;
    sum = block_reduce_add64<1024>(sum, ws64);
    if (threadIdx.x == 0)
        tile_counts[static_cast<size_t>(f) * max_tiles + blockIdx.x] = sum;
}

__global__ void __launch_bounds__(1024)
grid_scan_kernel(BatchView bv, TableView tv, const unsigned long long *__restrict__ tkeys,
                 uint32_t *__restrict__ tcount, uint4 *__restrict__ cells, uint32_t max_tiles,
                 const unsigned long long *__restrict__ tile_counts)
{
    __shared__ uint32_t ws[33];
    __shared__ unsigned long long ws64[33];
    const uint32_t f = blockIdx.y;
    const uint32_t cap = table_mask(bv.cnt[f], tv.tcap[f]) + 1u;
    const uint32_t toff = tv.toff[f];
    const uint32_t base = blockIdx.x * kScanTile;
    if (base >= cap)
        return;
    const uint32_t before =
        static_cast<uint32_t>(tile_prefix64<1024>(tile_counts + static_cast<size_t>(f) * max_tiles, blockIdx.x, ws64));
    const uint32_t first = base + threadIdx.x * kScanPer;
    uint32_t cnt[kScanPer];
    uint32_t sum = 0u;
#pragma unroll
    for (int k = 0; k < kScanPer; ++k)
    {
        cnt[k] = (first + k < cap) ? tcount[toff + first + k] : 0u;
        sum += cnt[k];
    }
    uint32_t total;
    uint32_t run = before + block_exclusive_scan<1024>(sum, ws, &total);
#pragma unroll
    for (int k = 0; k < kScanPer; ++k)
    {
        if (first + k < cap)
        {
            const unsigned long long key = tkeys[toff + first + k];
            cells[toff + first + k] = make_uint4(static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), run, cnt[k]);
            tcount[toff + first + k] = 0u;
        }
        run += cnt[k];
    }
}

// cpts[pos] = {x, y, z, bits(index)} grouped by cell; pos_of[index] = pos; cell_of[pos] = first pos of
// the point's cell — the id under which the cell is known to the union-find kernels
__global__ void __launch_bounds__(256)
grid_fill_kernel(const float4 *__restrict__ pts, BatchView bv, TableView tv, const uint4 *__restrict__ cells,
                 uint32_t *__restrict__ tcount, const uint32_t *__restrict__ slot_of, float4 *__restrict__ cpts,
                 uint32_t *__restrict__ pos_of, uint32_t *__restrict__ cell_of)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t toff = tv.toff[f];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    {
        const uint32_t slot = slot_of[off + i];
        const uint32_t first = cells[toff + slot].z;
        const uint32_t pos = first + atomicAdd(&tcount[toff + slot], 1u);
        const float4 p = pts[off + i];
        cpts[off + pos] = make_float4(p.x, p.y, p.z, __uint_as_float(i));
        pos_of[off + i] = pos;
        cell_of[off + pos] = first;
    }
}

// ---------------------------------------------------------------------------------------------
// lock-free union-find; roots are always the smallest id of their tree
LB_D uint32_t uf_find(uint32_t *parent, uint32_t x)
{
    uint32_t p = __ldcg(&parent[x]);
    while (p != x)
    {
        const uint32_t gp = __ldcg(&parent[p]);
        if (gp != p)
            parent[x] = gp; // path halving; only ever points to an ancestor
        x = p;
        p = gp;
    }
    return x;
}

LB_D void uf_unite(uint32_t *parent, uint32_t a, uint32_t b)
{
    while (true)
    {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b)
            return;
        if (a < b)
        {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        const uint32_t old = atomicCAS(&parent[a], a, b); // hook the larger root under the smaller
        if (old == a)
            return;
    }
}

__global__ void __launch_bounds__(256)
cc_init_kernel(BatchView bv, uint32_t *__restrict__ parent, uint32_t *__restrict__ comp_size)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    {
        parent[off + i] = i;
        comp_size[off + i] = 0u;
    }
}

// Union-find runs in CELL-ORDER space: parent[pos], roots are the smallest pos of their tree. A cell is
// a contiguous run of pos and is named by its first pos (cell id); per cell:
//   nbr[cid * 27 + k]  = id of the neighbour cell number k = (dz+1)*9 + (dy+1)*3 + (dx+1), or kNoCell
//   cinfo[cid]         = {points in the cell, common parent of all its points after the last
//                         flattening or kNoCell when they differ}
// The neighbour table is built once per frame (27 hash probes per CELL instead of per point).
constexpr uint32_t kNoCell = 0xFFFFFFFFu;

// One warp per 32 consecutive pos: for every cell that starts among them, lanes 0..26 probe the hash.
__global__ void __launch_bounds__(256)
cc_nbr_kernel(const float4 *__restrict__ cpts, BatchView bv, TableView tv, const uint4 *__restrict__ cells,
              const uint32_t *__restrict__ slot_of, const uint32_t *__restrict__ cell_of, uint32_t *__restrict__ nbr,
              uint2 *__restrict__ cinfo, uint32_t *__restrict__ nb27 /* packed copy for the replay: first pos | count << 20 */)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t mask = table_mask(m, tv.tcap[f]);
    const uint4 *tab = cells + tv.toff[f];
    const uint32_t lane = lane_id();
    const uint32_t warps_per_grid = gridDim.x * (blockDim.x >> 5);
    const long long dk = static_cast<long long>(static_cast<int>(lane % 3u) - 1) +
                         (static_cast<long long>(static_cast<int>((lane / 3u) % 3u) - 1) << 21) +
                         (static_cast<long long>(static_cast<int>(lane / 9u) - 1) << 42);
    for (uint32_t base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u; base < m; base += warps_per_grid * 32u)
    {
        const uint32_t pos = base + lane;
        uint32_t slot = 0u;
        bool is_start = false;
        if (pos < m)
        {
            is_start = cell_of[off + pos] == pos;
            if (is_start)
                slot = slot_of[off + __float_as_uint(cpts[off + pos].w)];
        }
        uint32_t starts = __ballot_sync(kFullMask, is_start);
        while (starts)
        {
            const int src = __ffs(starts) - 1;
            starts &= starts - 1u;
            const uint32_t cslot = __shfl_sync(kFullMask, slot, src);
            const uint32_t cid = base + static_cast<uint32_t>(src);
            const uint4 own = tab[cslot];
            if (lane < 27u)
            {
                const unsigned long long key =
                    (static_cast<unsigned long long>(own.x) | (static_cast<unsigned long long>(own.y) << 32)) +
                    static_cast<unsigned long long>(dk);
                uint32_t nstart, ncount;
                cell_lookup(tab, mask, key, &nstart, &ncount);
                nbr[(static_cast<size_t>(off) + cid) * 27u + lane] = ncount ? nstart : kNoCell;
                nb27[(static_cast<size_t>(off) + cid) * 27u + lane] =
                    ncount ? (nstart | ((ncount < 4095u ? ncount : 4095u) << 20)) : 0u;
            }
            if (lane == 0)
                cinfo[off + cid] = make_uint2(own.w, kNoCell);
        }
    }
}

// Afforest-style sampling pass, one thread per point: link to the first point of the own cell that is
// within reach (star-shaped trees), then to the first point within reach in each of the +x, +y, +z
// face neighbours. This already merges most of every component.
__global__ void __launch_bounds__(256)
cc_sample_kernel(const float4 *__restrict__ cpts, BatchView bv, CluParams prm, const uint32_t *__restrict__ cell_of,
                 const uint32_t *__restrict__ nbr, const uint2 *__restrict__ cinfo, uint32_t *__restrict__ parent)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const float4 *cp = cpts + off;
    uint32_t *par = parent + off;
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < m; pos += gridDim.x * blockDim.x)
    {
        const float4 pj = cp[pos];
        const uint32_t cid = cell_of[off + pos];
        for (uint32_t q = cid; q < pos; ++q)
        {
            const float4 cand = cp[q];
            if (dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z) <= prm.distance_squared)
            {
                uf_unite(par, pos, q);
                break;
            }
        }
        const uint32_t *row = nbr + (static_cast<size_t>(off) + cid) * 27u;
#pragma unroll
        for (int t = 0; t < 3; ++t)
        {
            const uint32_t n = row[t == 0 ? 14 : (t == 1 ? 16 : 22)];
            if (n == kNoCell)
                continue;
            const uint32_t cnt = cinfo[off + n].x;
            for (uint32_t q = n; q < n + cnt; ++q)
            {
                const float4 cand = cp[q];
                if (dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z) <= prm.distance_squared)
                {
                    uf_unite(par, pos, q);
                    break;
                }
            }
        }
    }
}

// cinfo[cid].y = the common parent of all points of the cell, or kNoCell when they differ. Run right
// after cc_compress_kernel. One thread per point, only cell starts work; cells are short.
__global__ void __launch_bounds__(256)
cc_cell_parent_kernel(BatchView bv, const uint32_t *__restrict__ cell_of, const uint32_t *__restrict__ parent,
                      uint2 *__restrict__ cinfo)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < m; pos += gridDim.x * blockDim.x)
    {
        if (cell_of[off + pos] != pos)
            continue;
        const uint32_t cnt = cinfo[off + pos].x;
        uint32_t common = parent[off + pos];
        for (uint32_t k = 1; k < cnt; ++k)
            if (parent[off + pos + k] != common)
            {
                common = kNoCell;
                break;
            }
        cinfo[off + pos].y = common;
    }
}

// Full pass, one thread per point. Every unordered pair of points in the same or in adjacent cells is
// visited exactly once: a point pairs with the points of its own cell that precede it and with the 13
// neighbour cells that follow its cell in (dz, dy, dx) lexicographic order. parent[] was flattened
// after the sampling pass, so nearly every neighbour cell carries this point's parent as its common
// parent and is skipped without touching a point; in the remaining cells a candidate whose (cached)
// parent equals the point's own is skipped before its coordinates are loaded. Stale cached parents
// are safe for that test: a node reachable through old pointers stays in the same set forever.
__global__ void __launch_bounds__(256)
cc_link_kernel(const float4 *__restrict__ cpts, BatchView bv, CluParams prm, const uint32_t *__restrict__ cell_of,
               const uint32_t *__restrict__ nbr, const uint2 *__restrict__ cinfo, uint32_t *__restrict__ parent)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const float4 *cp = cpts + off;
    uint32_t *par = parent + off;
    for (uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x; pos < m; pos += gridDim.x * blockDim.x)
    {
        const float4 pj = cp[pos];
        uint32_t ri = par[pos];
        const uint32_t cid = cell_of[off + pos];
        const uint32_t *row = nbr + (static_cast<size_t>(off) + cid) * 27u;
        for (uint32_t k = 13u; k < 27u; ++k)
        {
            const uint32_t n = k == 13u ? cid : row[k];
            if (n == kNoCell)
                continue;
            const uint2 ci = cinfo[off + n];
            if (ci.y == ri)
                continue;
            const uint32_t end = k == 13u ? pos : n + ci.x; // own cell: only the points that precede pos
            for (uint32_t q = n; q < end; ++q)
            {
                if (par[q] == ri)
                    continue;
                const float4 cand = cp[q];
                if (dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z) <= prm.distance_squared)
                {
                    uf_unite(par, pos, q);
                    ri = __ldcg(&par[pos]);
                }
            }
        }
    }
}

// parent[pos] = root(pos), in place
__global__ void __launch_bounds__(256) cc_compress_kernel(BatchView bv, uint32_t *__restrict__ parent)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    {
        uint32_t x = i;
        uint32_t p = __ldcg(&parent[off + x]);
        while (p != x)
        {
            x = p;
            p = __ldcg(&parent[off + x]);
        }
        parent[off + i] = x;
    }
}

// For point index i: keys[i] = component id (root pos) of i, vals[i] = i; comp_size[root pos] = members.
// A stable sort by key then lists every component's members in ascending index order.
__global__ void __launch_bounds__(256)
cc_flatten_kernel(BatchView bv, const uint32_t *__restrict__ parent, const uint32_t *__restrict__ pos_of,
                  uint32_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ comp_size)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    {
        uint32_t x = pos_of[off + i];
        uint32_t p = __ldcg(&parent[off + x]);
        while (p != x)
        {
            x = p;
            p = __ldcg(&parent[off + x]);
        }
        keys[off + i] = x;
        vals[off + i] = i;
        // warp-aggregated count per root (neighbouring indices mostly share a component)
        const uint32_t peers = __match_any_sync(__activemask(), x);
        if ((peers & lanemask_lt()) == 0u)
            atomicAdd(&comp_size[off + x], static_cast<uint32_t>(__popc(peers)));
    }
}

// Replay working set, in cell order: rpts[pos] = {x, y, z, bits(state)} with state = rank << 2 | flags,
// so one 16-byte load brings a candidate's coordinates, its k-d pre-order rank and its removed /
// queued flags. pkey[pos] = key of the point's cell (neighbour keys are one 64-bit add away). seed_of[pos] = unset; member_pos[t] = pos of the t-th member; pslot[pos] = hash slot
// of the point's cell; tlive[slot] = points of the cell that are not removed yet.
__global__ void __launch_bounds__(256)
replay_init_kernel(const float4 *__restrict__ cpts, BatchView bv, const uint32_t *__restrict__ rank_of_point,
                   const uint32_t *__restrict__ member_idx, const uint32_t *__restrict__ pos_of,
                   const uint32_t *__restrict__ slot_of, float4 *__restrict__ rpts, uint32_t *__restrict__ seed_of,
                   uint32_t *__restrict__ member_pos, uint32_t *__restrict__ pslot, uint32_t *__restrict__ cursor,
                   TableView tv, const uint4 *__restrict__ cells, unsigned long long *__restrict__ pkey)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 16)
        cursor[threadIdx.x] = 0u; // [0,1] warp-path phases, [2] CTA-path cursor, [3..6] CTA-path job counts per size bucket,
                                  // [7,8] cursor / count of the huge list, [9,10] of the first-generation list, [11..13] stay 0 (that list is read as four size buckets), [14] cursor of the short-job launch of the window-synchronous replay
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    {
        const float4 p = cpts[off + i];
        const uint32_t idx = __float_as_uint(p.w);
        rpts[off + i] = make_float4(p.x, p.y, p.z, __uint_as_float(rank_of_point[off + idx] << 2));
        seed_of[off + i] = kSeedUnset;
        const uint32_t slot = slot_of[off + idx];
        pslot[off + i] = slot;
        const uint4 cell = cells[tv.toff[f] + slot];
        pkey[off + i] = static_cast<unsigned long long>(cell.x) | (static_cast<unsigned long long>(cell.y) << 32);
        member_pos[off + i] = pos_of[off + member_idx[off + i]];
    }
}

// tlive[slot] = cells[slot].count
__global__ void __launch_bounds__(256)
replay_live_init_kernel(BatchView bv, TableView tv, const uint4 *__restrict__ cells, uint32_t *__restrict__ tlive)
{
    const uint32_t f = blockIdx.y;
    const uint32_t cap = table_mask(bv.cnt[f], tv.tcap[f]) + 1u;
    const uint32_t toff = tv.toff[f];
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < cap; s += gridDim.x * blockDim.x)
        tlive[toff + s] = cells[toff + s].w;
}

constexpr int kReplayWarps = 4;
constexpr uint32_t kPushCap = 256u;     // per-warp shared push buffer; larger expansions spill to global
constexpr uint32_t kBigComponent = 96u; // components with at least this many members are replayed first
constexpr uint32_t kCtaComponentMin = 128u; // smallest permitted CluParams::cta_min_members (sizes the job list)
constexpr int kReplayUnroll = 4;        // candidate batches whose loads are issued together

// Replay-time lookup: tlive[slot] counts the points of the cell that are not removed yet (decremented
// with fire-and-forget atomics by whoever removes a point); a cell without live points cannot
// contribute (clustering.cpp:94-97) and reports count 0. A stale non-zero only costs a wasted scan.
LB_D void cell_lookup_alive(const uint4 *cells, const uint32_t *tlive, uint32_t mask, uint64_t key, uint32_t *start,
                            uint32_t *count, uint32_t *slot_out)
{
    uint32_t slot = hash_cell(key) & mask;
    const uint32_t klo = static_cast<uint32_t>(key), khi = static_cast<uint32_t>(key >> 32);
    while (true)
    {
        const uint4 c = __ldg(&cells[slot]);
        const uint32_t live = __ldcg(&tlive[slot]);
        if (c.x == klo && c.y == khi)
        {
            *start = c.z;
            *count = live ? c.w : 0u;
            *slot_out = slot;
            return;
        }
        if (c.x == 0xFFFFFFFFu && c.y == 0xFFFFFFFFu)
        {
            *start = 0u;
            *count = 0u;
            *slot_out = 0u;
            return;
        }
        slot = (slot + 1u) & mask;
    }
}

// Persistent warps over a flat work list of (frame, 32 member slots) claims, walked twice: the first
// walk replays only the big components — their BFS chains are the critical path, so they must start
// at time zero — the second walk replays everything else around them. A warp replays every
// component whose first member falls into its claim.
__global__ void __launch_bounds__(kReplayWarps * 32)
replay_kernel(float4 *__restrict__ rpts_all, BatchView bv, TableView tv, const uint4 *__restrict__ cells,
              CluParams prm, const uint32_t *__restrict__ member_root, const uint32_t *__restrict__ member_idx,
              const uint32_t *__restrict__ member_pos, const uint32_t *__restrict__ comp_size,
              const uint32_t *__restrict__ pslot_all, uint32_t *__restrict__ tlive_all, uint32_t *__restrict__ seed_of,
              uint32_t *__restrict__ queue, unsigned long long *__restrict__ push_spill,
              uint8_t *__restrict__ seed_valid, uint32_t *__restrict__ cursor, uint32_t claims_per_frame)
{
    __shared__ unsigned long long pbuf_all[kReplayWarps][kPushCap];
    const uint32_t lane = lane_id();
    const uint32_t lt = lanemask_lt();
    unsigned long long *pbuf = pbuf_all[threadIdx.x >> 5];
    const uint32_t total_claims = claims_per_frame * bv.frames;

    for (uint32_t phase = 0; phase < 2u; ++phase)
    while (true)
    {
        uint32_t w = 0u;
        if (lane == 0)
            w = atomicAdd(&cursor[phase], 1u);
        w = __shfl_sync(kFullMask, w, 0);
        if (w >= total_claims)
            break;
        const uint32_t f = w / claims_per_frame;
        const uint32_t t0 = (w - f * claims_per_frame) * 32u;
        const uint32_t m = bv.cnt[f];
        if (t0 >= m)
            continue;
        const uint32_t off = bv.off[f];
        const uint32_t mask = table_mask(m, tv.tcap[f]);
        const uint4 *tab = cells + tv.toff[f];
        uint32_t *tlive = tlive_all + tv.toff[f];
        float4 *rp = rpts_all + off;
        // the state word of a point is the .w lane of its float4
        uint32_t *stw = reinterpret_cast<uint32_t *>(rp) + 3;
        const uint32_t *pslot = pslot_all + off;
        uint32_t *so = seed_of + off;
        uint32_t *qu = queue + off;
        unsigned long long *spill = push_spill + off;
        const uint32_t *mroot = member_root + off;
        const uint32_t *midx = member_idx + off;
        const uint32_t *mpos = member_pos + off;

        const uint32_t tt = t0 + lane;
        bool is_start = tt < m && (tt == 0u || mroot[tt] != mroot[tt - 1u]);
        if (is_start)
        {
            const uint32_t size = comp_size[off + mroot[tt]];
            is_start = size < prm.cta_min_members && (size >= kBigComponent) == (phase == 0u);
        }
        uint32_t starts = __ballot_sync(kFullMask, is_start);
        while (starts)
        {
            const uint32_t t_start = t0 + (__ffs(starts) - 1);
            starts &= starts - 1u;
            const uint32_t root = mroot[t_start];
            const uint32_t qbase = t_start; // the component's FIFO lives in queue[t_start, t_start + size)

            uint32_t u = t_start; // next member to examine as a seed candidate (ascending index)
            while (true)
            {
                // find the next member that is not removed (clustering.cpp:70-75)
                uint32_t seed_t = 0xFFFFFFFFu;
                bool component_done = false;
                while (true)
                {
                    const uint32_t uu = u + lane;
                    const bool in_comp = uu < m && mroot[uu] == root;
                    bool cand = false;
                    if (in_comp)
                        cand = (stw[4u * mpos[uu]] & kStRemoved) == 0u;
                    const uint32_t bc = __ballot_sync(kFullMask, cand);
                    const uint32_t bi = __ballot_sync(kFullMask, in_comp);
                    if (bc)
                    {
                        seed_t = u + (__ffs(bc) - 1);
                        break;
                    }
                    if (bi != kFullMask)
                    {
                        component_done = true;
                        break;
                    }
                    u += 32u;
                }
                if (component_done)
                    break;
                u = seed_t + 1u;
                const uint32_t seed_idx = midx[seed_t];
                const uint32_t seed_pos = mpos[seed_t];

                uint32_t head = 0u, tail = 0u, touched = 0u;
                if (lane == 0)
                {
                    qu[qbase] = seed_pos;
                    stw[4u * seed_pos] |= kStQueued;
                }
                tail = 1u;
                __syncwarp();

                while (head < tail) // clustering.cpp:80-111
                {
                    // Pop: look at the next 32 FIFO entries at once. Entries that are already removed
                    // are no-ops in the reference (clustering.cpp:85-88) and nothing between here and
                    // the first live entry can change any state, so jump straight to it.
                    const uint32_t e_look = head + lane;
                    uint32_t qj = 0u;
                    float4 pe = make_float4(0.f, 0.f, 0.f, __uint_as_float(kStRemoved));
                    if (e_look < tail)
                    {
                        qj = qu[qbase + e_look];
                        pe = rp[qj];
                    }
                    const uint32_t alive = __ballot_sync(kFullMask, (__float_as_uint(pe.w) & kStRemoved) == 0u);
                    if (alive == 0u)
                    {
                        head = min(tail, head + 32u);
                        continue;
                    }
                    const int first = __ffs(alive) - 1;
                    head += static_cast<uint32_t>(first) + 1u;
                    float4 pj;
                    pj.x = __shfl_sync(kFullMask, pe.x, first);
                    pj.y = __shfl_sync(kFullMask, pe.y, first);
                    pj.z = __shfl_sync(kFullMask, pe.z, first);
                    pj.w = 0.f;

                    int cx, cy, cz;
                    cell_coords(pj, prm.inv_cell, &cx, &cy, &cz);
                    uint32_t start = 0u, count = 0u;
                    if (lane < 27u)
                    {
                        uint32_t slot_unused;
                        cell_lookup_alive(tab, tlive, mask,
                                          cell_key(cx + static_cast<int>(lane % 3u) - 1,
                                                   cy + static_cast<int>((lane / 3u) % 3u) - 1,
                                                   cz + static_cast<int>(lane / 9u) - 1),
                                          &start, &count, &slot_unused);
                    }
                    const uint32_t incl = warp_inclusive_scan(count);
                    const uint32_t excl = incl - count;
                    const uint32_t total = __shfl_sync(kFullMask, incl, 31);
                    uint32_t np = 0u;
                    bool spilled = false;
                    for (uint32_t base = 0; base < total; base += 32u * kReplayUnroll)
                    {
                        // several batches of 32 candidates per trip: their loads are in flight together
                        uint32_t pos2[kReplayUnroll];
                        float4 cand2[kReplayUnroll];
                        bool valid2[kReplayUnroll];
#pragma unroll
                        for (int h = 0; h < kReplayUnroll; ++h)
                        {
                            const uint32_t q = base + 32u * h + lane;
                            uint32_t lo = 0u, hi = 26u;
#pragma unroll
                            for (int it = 0; it < 5; ++it)
                            {
                                const uint32_t mid = (lo + hi) >> 1;
                                const uint32_t v = __shfl_sync(kFullMask, incl, mid);
                                if (v > q)
                                    hi = mid;
                                else
                                    lo = mid + 1u;
                            }
                            const uint32_t c = min(lo, 26u);
                            const uint32_t cstart = __shfl_sync(kFullMask, start, c);
                            const uint32_t cexcl = __shfl_sync(kFullMask, excl, c);
                            valid2[h] = q < total;
                            pos2[h] = cstart + (q - cexcl);
                            cand2[h] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kStRemoved));
                            if (valid2[h])
                                cand2[h] = rp[pos2[h]];
                        }
#pragma unroll
                        for (int h = 0; h < kReplayUnroll; ++h)
                        {
                            if (h > 0 && base + 32u * h >= total)
                                break;
                            const uint32_t pos = pos2[h];
                            const float4 cand = cand2[h];
                            const uint32_t sw = __float_as_uint(cand.w);
                            bool live = false, push = false, removed_now = false;
                            if (valid2[h] && (sw & kStRemoved) == 0u) // removed points are skipped (clustering.cpp:94-97)
                            {
                                // KDTree::dist_sqr(target, node) (kdtree.hpp:145-163), inclusive test (kdtree.hpp:314)
                                const float d2 = dist_sqr_ref(pj.x, pj.y, pj.z, cand.x, cand.y, cand.z);
                                live = d2 <= prm.distance_squared;
                                if (live)
                                {
                                    so[pos] = seed_idx; // labels[k] = label (clustering.cpp:99)
                                    if (d2 <= prm.inner_threshold)
                                    {
                                        stw[4u * pos] = sw | kStRemoved; // clustering.cpp:102-105
                                        removed_now = true;
                                    }
                                    else if ((sw & kStQueued) == 0u)
                                    {
                                        stw[4u * pos] = sw | kStQueued; // clustering.cpp:106-109 (first push only)
                                        push = true;
                                    }
                                }
                            }
                            touched += __popc(__ballot_sync(kFullMask, live)); // indices_.push_back (with multiplicity)
                            if (__any_sync(kFullMask, removed_now))
                            {
                                // one fire-and-forget atomic per cell: candidates arrive cell by cell, so equal slots are neighbours
                                const uint32_t sl = removed_now ? pslot[pos] : 0xFFFFFFFFu;
                                const uint32_t peers = __match_any_sync(kFullMask, sl);
                                if (removed_now && (peers & lt) == 0u)
                                    atomicSub(&tlive[sl], static_cast<uint32_t>(__popc(peers)));
                            }
                            const uint32_t bp = __ballot_sync(kFullMask, push);
                            const uint32_t nadd = __popc(bp);
                            if (nadd)
                            {
                                if (!spilled && np + nadd > kPushCap)
                                {
                                    for (uint32_t e = lane; e < np; e += 32u)
                                        spill[qbase + tail + e] = pbuf[e];
                                    spilled = true;
                                    __syncwarp();
                                }
                                if (push)
                                {
                                    const unsigned long long ent = (static_cast<unsigned long long>(sw >> 2) << 32) |
                                                                   static_cast<unsigned long long>(pos);
                                    const uint32_t e = np + __popc(bp & lt);
                                    if (spilled)
                                        spill[qbase + tail + e] = ent;
                                    else
                                        pbuf[e] = ent;
                                }
                                np += nadd;
                            }
                        }
                    }
                    __syncwarp();
                    // the FIFO receives this expansion's pushes in ascending k-d pre-order rank
                    if (np)
                    {
                        if (!spilled && np > 32u)
                        {
                            // bitonic sort in the warp's shared buffer
                            uint32_t n2 = 64u;
                            while (n2 < np)
                                n2 <<= 1;
                            for (uint32_t e = np + lane; e < n2; e += 32u)
                                pbuf[e] = 0xFFFFFFFFFFFFFFFFull;
                            __syncwarp();
                            for (uint32_t kk = 2u; kk <= n2; kk <<= 1)
                                for (uint32_t jj = kk >> 1; jj > 0u; jj >>= 1)
                                {
                                    for (uint32_t t = lane; t < (n2 >> 1); t += 32u)
                                    {
                                        const uint32_t i0 = ((t & ~(jj - 1u)) << 1) | (t & (jj - 1u));
                                        const uint32_t i1 = i0 | jj;
                                        const unsigned long long a = pbuf[i0], b = pbuf[i1];
                                        const bool asc = (i0 & kk) == 0u;
                                        if ((a > b) == asc)
                                        {
                                            pbuf[i0] = b;
                                            pbuf[i1] = a;
                                        }
                                    }
                                    __syncwarp();
                                }
                            for (uint32_t e = lane; e < np; e += 32u)
                                qu[qbase + tail + e] = static_cast<uint32_t>(pbuf[e]);
                        }
                        else
                        {
                            const unsigned long long *src = spilled ? (spill + qbase + tail) : pbuf;
                            for (uint32_t e = lane; e < np; e += 32u)
                            {
                                const unsigned long long mine = src[e];
                                uint32_t dest = 0u;
                                for (uint32_t x = 0; x < np; ++x)
                                    dest += (src[x] < mine) ? 1u : 0u;
                                qu[qbase + tail + dest] = static_cast<uint32_t>(mine);
                            }
                        }
                        tail += np;
                    }
                    __syncwarp();
                }
                if (lane == 0) // clustering.cpp:113-123
                    seed_valid[off + seed_idx] =
                        (touched < prm.min_cluster_size || touched > prm.max_cluster_size) ? 0u : 1u;
            }
        }
    }
}

// label k = number of valid seeds with a smaller index (clustering.cpp:68,113-123), in tiles of 4096
// points (grid = (tiles, frames)): count the valid seeds per tile, number them, then label every point.
constexpr int kLabelPer = 4;
constexpr uint32_t kLabelTile = 1024u * kLabelPer;

LB_D bool label_is_valid_seed(uint32_t i, uint32_t off, const uint32_t *pos_of, const uint32_t *seed_of,
                              const uint8_t *seed_valid)
{
    return seed_of[off + pos_of[off + i]] == i && seed_valid[off + i] != 0u;
}

__global__ void __launch_bounds__(1024)
label_count_kernel(BatchView bv, const uint32_t *__restrict__ pos_of, const uint32_t *__restrict__ seed_of,
                   const uint8_t *__restrict__ seed_valid, uint32_t max_tiles, unsigned long long *__restrict__ tile_counts)
{
    __shared__ unsigned long long ws64[33];
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t first = blockIdx.x * kLabelTile + threadIdx.x * kLabelPer;
    unsigned long long sum = 0ull;
#pragma unroll
    for (int k = 0; k < kLabelPer; ++k)
        if (first + k < m && label_is_valid_seed(first + k, off, pos_of, seed_of, seed_valid))
            ++sum;
    sum = block_reduce_add64<1024>(sum, ws64);
    if (threadIdx.x == 0)
        tile_counts[static_cast<size_t>(f) * max_tiles + blockIdx.x] = sum;
}

__global__ void __launch_bounds__(1024)
label_number_kernel(BatchView bv, const uint32_t *__restrict__ pos_of, const uint32_t *__restrict__ seed_of,
                    const uint8_t *__restrict__ seed_valid, uint32_t max_tiles,
                    const unsigned long long *__restrict__ tile_counts, uint32_t *__restrict__ seed_label,
                    uint32_t *__restrict__ n_clusters)
{
    __shared__ uint32_t ws[33];
    __shared__ unsigned long long ws64[33];
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    const uint32_t base = blockIdx.x * kLabelTile;
    if (base >= m && blockIdx.x != 0u)
        return;
    const uint32_t before =
        static_cast<uint32_t>(tile_prefix64<1024>(tile_counts + static_cast<size_t>(f) * max_tiles, blockIdx.x, ws64));
    const uint32_t first = base + threadIdx.x * kLabelPer;
    bool is_seed[kLabelPer];
    uint32_t sum = 0u;
#pragma unroll
    for (int k = 0; k < kLabelPer; ++k)
    {
        is_seed[k] = first + k < m && label_is_valid_seed(first + k, off, pos_of, seed_of, seed_valid);
        sum += is_seed[k] ? 1u : 0u;
    }
    uint32_t total;
    uint32_t run = before + block_exclusive_scan<1024>(sum, ws, &total);
#pragma unroll
    for (int k = 0; k < kLabelPer; ++k)
        if (is_seed[k])
            seed_label[off + first + k] = run++;
    if (threadIdx.x == 0 && base + kLabelTile >= m)
        n_clusters[f] = before + total;
}

__global__ void __launch_bounds__(256)
label_assign_kernel(BatchView bv, const uint32_t *__restrict__ pos_of, const uint32_t *__restrict__ seed_of,
                    const uint8_t *__restrict__ seed_valid, const uint32_t *__restrict__ seed_label,
                    int32_t *__restrict__ labels)
{
    const uint32_t f = blockIdx.y;
    const uint32_t m = bv.cnt[f];
    const uint32_t off = bv.off[f];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    {
        const uint32_t s = seed_of[off + pos_of[off + i]];
        int32_t lab = kLabelUndefined;
        if (s != kSeedUnset)
            lab = seed_valid[off + s] ? static_cast<int32_t>(seed_label[off + s]) : kLabelInvalid;
        labels[off + i] = lab;
    }
}

} // namespace lb
