// Host-side checks of the PRODUCT headers that are host-compilable (kd_select.h, jacobi3.h, chi_shape.h) and a
// lane-by-lane model of the cooperative Hoare-partition formulation used by kd_build.cuh.
// Built by __graft_entry__.build() into tests/host/libhost_checks.so; driven by tests/test_host_logic.py.
#include "../../lidar-processing_b200/csrc/chi_shape.h"
#include "../../lidar-processing_b200/csrc/jacobi3.h"
#include "../../lidar-processing_b200/csrc/kd_select.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

using lb::kd_key;

namespace
{
// One cooperative partition round over [first+1,last) around a[first], evaluated the way the CUDA
// kernel does: NW warps own contiguous 32-aligned chunks, lanes interleave, ranks come from ballots.
// Returns the cut.
uint32_t model_partition_round(float4 *a, uint32_t first, uint32_t last, int axis, uint32_t nw,
                               std::vector<uint32_t> &gepos, std::vector<uint32_t> &lepos)
{
    const float kp = kd_key(a[first], axis);
    const uint32_t rb = first + 1u;
    const uint32_t r = last - rb;
    uint32_t chunk = (r + nw - 1u) / nw;
    chunk = (chunk + 31u) / 32u * 32u;
    std::vector<uint32_t> ge_w(nw, 0u), le_w(nw, 0u);
    for (uint32_t w = 0; w < nw; ++w)
    {
        const uint32_t cb = rb + w * chunk;
        const uint32_t ce = std::min(last, cb + chunk);
        for (uint32_t p = cb; p < ce && cb < last; ++p)
        {
            const float k = kd_key(a[p], axis);
            ge_w[w] += !(k < kp);
            le_w[w] += !(kp < k);
        }
    }
    uint32_t K = 0u, min_unswapped_ge = 0xFFFFFFFFu, min_swapped_le = 0xFFFFFFFFu;
    for (uint32_t w = 0; w < nw; ++w)
    {
        uint32_t ge_before = 0u, le_after = 0u;
        for (uint32_t v = 0; v < w; ++v)
            ge_before += ge_w[v];
        for (uint32_t v = w + 1; v < nw; ++v)
            le_after += le_w[v];
        const uint32_t cb = rb + w * chunk;
        if (cb >= last)
            continue;
        const uint32_t ce = std::min(last, cb + chunk);
        uint32_t run_ge = 0u, run_le = 0u;
        for (uint32_t itb = cb; itb < ce; itb += 32u)
        {
            uint32_t bge = 0u, ble = 0u;
            for (uint32_t lane = 0; lane < 32u; ++lane)
            {
                const uint32_t p = itb + lane;
                if (p < ce)
                {
                    const float k = kd_key(a[p], axis);
                    if (!(k < kp))
                        bge |= 1u << lane;
                    if (!(kp < k))
                        ble |= 1u << lane;
                }
            }
            for (uint32_t lane = 0; lane < 32u; ++lane)
            {
                const uint32_t p = itb + lane;
                if (p >= ce)
                    continue;
                const uint32_t lt = (1u << lane) - 1u;
                const uint32_t le_incl = lt | (1u << lane);
                const uint32_t ge_left = ge_before + run_ge + __builtin_popcount(bge & lt);
                const uint32_t le_right = le_after + (le_w[w] - run_le - __builtin_popcount(ble & le_incl));
                const bool is_ge = (bge >> lane) & 1u, is_le = (ble >> lane) & 1u;
                if (is_ge)
                {
                    if (le_right > ge_left)
                    {
                        gepos[rb + ge_left] = p;
                        ++K;
                    }
                    else
                        min_unswapped_ge = std::min(min_unswapped_ge, p);
                }
                if (is_le && ge_left > le_right)
                {
                    lepos[rb + le_right] = p;
                    min_swapped_le = std::min(min_swapped_le, p);
                }
            }
            run_ge += __builtin_popcount(bge);
            run_le += __builtin_popcount(ble);
        }
    }
    for (uint32_t k = 0; k < K; ++k)
        lb::kd_swap(a, gepos[rb + k], lepos[rb + k]);
    return std::min(min_unswapped_ge, min_swapped_le);
}

void model_coop_nth_element(float4 *a, uint32_t first, uint32_t nth, uint32_t last, int axis, uint32_t nw,
                            uint32_t seq_cutoff, std::vector<uint32_t> &gepos, std::vector<uint32_t> &lepos)
{
    if (first == last || nth == last)
        return;
    uint32_t depth_limit = 2u * lb::kd_floor_log2(last - first);
    while (last - first > 3u && last - first > seq_cutoff && depth_limit != 0u)
    {
        --depth_limit;
        const uint32_t mid = first + (last - first) / 2u;
        lb::kd_move_median_to_first(a, first, first + 1u, mid, last - 1u, axis);
        const uint32_t cut = model_partition_round(a, first, last, axis, nw, gepos, lepos);
        if (cut <= nth)
            first = cut;
        else
            last = cut;
    }
    lb::kd_introselect_from(a, first, nth, last, depth_limit, axis);
}
} // namespace

// rec[i] = {rank of dist[i] among the distinct distances, i}: what the device builds from its sorted keys
void rank_records(const double *dist, uint32_t n, lb::ChiKeyed *rec)
{
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; ++i)
        order[i] = i;
    std::sort(order.begin(), order.end(), [dist](uint32_t x, uint32_t y) { return dist[x] < dist[y] || (dist[x] == dist[y] && x < y); });
    uint32_t rank = 0u;
    for (uint32_t j = 0; j < n; ++j)
    {
        if (j > 0u && dist[order[j]] != dist[order[j - 1u]])
            ++rank;
        rec[order[j]].d = rank;
        rec[order[j]].id = order[j];
    }
}

extern "C"
{

// mode 0: std::nth_element; 1: lb::kd_nth_element (sequential product code); 2: cooperative model
void hc_nth_element(float4 *a, uint32_t first, uint32_t nth, uint32_t last, int axis, int mode, uint32_t nw,
                    uint32_t seq_cutoff)
{
    if (mode == 0)
        std::nth_element(a + first, a + nth, a + last,
                         [axis](const float4 &x, const float4 &y) { return kd_key(x, axis) < kd_key(y, axis); });
    else if (mode == 1)
        lb::kd_nth_element(a, first, nth, last, axis);
    else
    {
        std::vector<uint32_t> ge(last + 1u), le(last + 1u);
        model_coop_nth_element(a, first, nth, last, axis, nw, seq_cutoff, ge, le);
    }
}

// Builds the implicit k-d tree in `a` (m nodes) and writes rank_out[index] = pre-order rank.
void hc_kd_build(float4 *a, uint32_t m, int mode, uint32_t nw, uint32_t seq_cutoff, uint32_t *rank_out)
{
    struct R
    {
        uint32_t b, e, d;
    };
    std::vector<R> st;
    std::vector<uint32_t> ge(m + 1u), le(m + 1u);
    st.push_back({0u, m, 0u});
    while (!st.empty())
    {
        const R r = st.back();
        st.pop_back();
        if (r.b >= r.e)
            continue;
        const uint32_t mid = r.b + (r.e - r.b) / 2u;
        const int axis = static_cast<int>(r.d % 3u);
        if (mode == 0)
            std::nth_element(a + r.b, a + mid, a + r.e,
                             [axis](const float4 &x, const float4 &y) { return kd_key(x, axis) < kd_key(y, axis); });
        else if (mode == 1)
            lb::kd_nth_element(a, r.b, mid, r.e, axis);
        else
            model_coop_nth_element(a, r.b, mid, r.e, axis, nw, seq_cutoff, ge, le);
        st.push_back({r.b, mid, r.d + 1u});
        st.push_back({mid + 1u, r.e, r.d + 1u});
    }
    for (uint32_t s = 0; s < m; ++s)
    {
        uint32_t idx;
        std::memcpy(&idx, &a[s].w, 4);
        rank_out[idx] = lb::kd_preorder_rank_of_slot(m, s);
    }
}

int hc_range_at(uint32_t m, uint32_t depth, uint32_t path, uint32_t *b, uint32_t *e)
{
    return lb::kd_range_at(m, depth, path, b, e) ? 1 : 0;
}

int hc_jacobi_svd3(const float *a, float *v, float *sv)
{
    return lb::jacobi_svd3(a, v, sv) ? 1 : 0;
}

// bytes of the working set of a cluster of n points (the device places it at 96 bytes per point)
unsigned long long hc_chi_layout_bytes(uint32_t n)
{
    return lb::chi_layout(n).bytes;
}

// distances from the seed circumcentre of one cluster as the sweep sorts them (diagnostics); returns the seed status
uint32_t hc_chi_dists(const float *points, uint32_t n, uint32_t stride_floats, double *dist_out)
{
    const lb::ChiLayout lay = lb::chi_layout(n);
    std::vector<unsigned char> block(lay.bytes, 0);
    lb::ChiWork w;
    lb::chi_bind(w, block.data(), lay, n);
    for (uint32_t i = 0; i < n; ++i)
    {
        w.node[i].x = static_cast<double>(points[static_cast<size_t>(i) * stride_floats]);
        w.node[i].y = static_cast<double>(points[static_cast<size_t>(i) * stride_floats + 1]);
    }
    const uint32_t err = lb::chi_seed_sequential(w);
    if (err == lb::kChiOk)
        std::memcpy(dist_out, w.dist, sizeof(double) * n);
    return err;
}

// std::sort of ids by dist (mode 0) against lb::chi_introsort_ids (mode 1), for the tie-order check
void hc_sort_ids(uint32_t *ids, const double *dist, uint32_t n, int mode)
{
    if (mode == 0)
        std::sort(ids, ids + n, [dist](uint32_t i, uint32_t j) { return dist[i] < dist[j]; });
    else
    {
        std::vector<lb::ChiKeyed> rec(n + 1u);
        rank_records(dist, n, rec.data());
        lb::chi_introsort(rec.data(), n);
        for (uint32_t i = 0; i < n; ++i)
            ids[i] = rec[i].id;
    }
}

// std::push_heap / std::pop_heap on (edge, length) pairs against lb::chi_heap_push / chi_heap_pop: ops[i] >= 0 pushes
// lens[i] with edge number i, ops[i] < 0 pops; popped_out receives the edge numbers in pop order. Returns their count.
uint32_t hc_heap_replay(const int *ops, const double *lens, uint32_t n_ops, int mode, uint32_t *popped_out)
{
    uint32_t n_pop = 0u;
    if (mode == 0)
    {
        using HP = std::pair<std::size_t, double>;
        const auto cmp = [](const HP &l, const HP &r) { return l.second < r.second; };
        std::vector<HP> h;
        for (uint32_t i = 0; i < n_ops; ++i)
        {
            if (ops[i] >= 0)
            {
                h.emplace_back(i, lens[i]);
                std::push_heap(h.begin(), h.end(), cmp);
            }
            else if (!h.empty())
            {
                std::pop_heap(h.begin(), h.end(), cmp);
                popped_out[n_pop++] = static_cast<uint32_t>(h.back().first);
                h.pop_back();
            }
        }
    }
    else
    {
        std::vector<uint32_t> he(n_ops + 1u);
        std::vector<double> hl(n_ops + 1u);
        uint32_t size = 0u;
        for (uint32_t i = 0; i < n_ops; ++i)
        {
            if (ops[i] >= 0)
                lb::chi_heap_push(he.data(), hl.data(), size, i, lens[i]);
            else if (size)
            {
                uint32_t e;
                double l;
                lb::chi_heap_pop(he.data(), hl.data(), size, e, l);
                popped_out[n_pop++] = e;
            }
        }
    }
    return n_pop;
}

// Concave outlines of CSR clusters with the product's sequential core (chi_shape.h), the way chi_shape.cuh drives it.
// sort_mode 0: order by (distance, index), std::sort re-enactment only when two different points are exactly equally
// far (the device's policy); 1: always the re-enactment; 2: always (distance, index).
// sizes_out[k]: vertices of the closed outline (0: fewer than 20 points, not this function's business; 0xFFFFFFFF: the
// reference throws / reads out of bounds). idx_out: cluster-local vertex indices end to end. stats_out[0] = clusters
// that needed the re-enactment, [1] = clusters run. Returns the number of indices written or -1 (capacity).
long long hc_chi_outlines(const float *points, const uint32_t *offsets, uint32_t n_clusters, uint32_t stride_floats,
                          int sort_mode, uint32_t *sizes_out, uint32_t *idx_out, long long capacity, uint32_t *stats_out)
{
    long long total = 0;
    std::vector<unsigned char> block;
    std::vector<uint32_t> loop;
    std::vector<std::pair<double, uint32_t>> order;
    stats_out[0] = stats_out[1] = 0u;
    for (uint32_t k = 0; k < n_clusters; ++k)
    {
        const uint32_t n = offsets[k + 1] - offsets[k];
        sizes_out[k] = 0u;
        if (n < 20u)
            continue;
        ++stats_out[1];
        const lb::ChiLayout lay = lb::chi_layout(n);
        block.assign(lay.bytes, 0xCD);
        lb::ChiWork w;
        lb::chi_bind(w, block.data(), lay, n);
        for (uint32_t i = 0; i < n; ++i)
        {
            const float *p = points + static_cast<size_t>(offsets[k] + i) * stride_floats;
            w.node[i].x = static_cast<double>(p[0]);
            w.node[i].y = static_cast<double>(p[1]);
        }
        uint32_t err = lb::chi_seed_sequential(w);
        if (err == lb::kChiOk)
        {
            order.resize(n);
            for (uint32_t i = 0; i < n; ++i)
                order[i] = {w.dist[i], i};
            std::sort(order.begin(), order.end());
            bool mixed = false;
            for (uint32_t i = 0; i + 1 < n; ++i)
                if (order[i].first == order[i + 1].first)
                {
                    // equally far and not the same place - unless both lie on seed vertices (all of those are skipped
                    // whatever their order, and the three seeds are equally far from their circumcentre by construction)
                    const lb::ChiNode &a = w.node[order[i].second], &b = w.node[order[i + 1].second];
                    if (!(a.x == b.x && a.y == b.y) && !(lb::chi_on_seed(w, a.x, a.y) && lb::chi_on_seed(w, b.x, b.y)))
                        mixed = true;
                }
            if (sort_mode == 1 || (sort_mode == 0 && mixed))
            {
                lb::ChiKeyed *rec = reinterpret_cast<lb::ChiKeyed *>(w.edge); // (the device's scratch too)
                rank_records(w.dist, n, rec);
                lb::chi_introsort(rec, n);
                for (uint32_t i = 0; i < n; ++i)
                    w.ids[i] = rec[i].id;
                ++stats_out[0];
            }
            else
                for (uint32_t i = 0; i < n; ++i)
                    w.ids[i] = order[i].second;
            err = lb::chi_triangulate(w);
        }
        if (err != lb::kChiOk)
        {
            sizes_out[k] = 0xFFFFFFFFu;
            continue;
        }
        loop.resize(n);
        const uint32_t h = lb::chi_erode_and_walk(w, loop.data());
        if (total + h > capacity)
            return -1;
        sizes_out[k] = h;
        for (uint32_t v = 0; v + 1 < h; ++v)
            idx_out[total++] = loop[v];
        idx_out[total++] = loop[0];
    }
    return total;
}

} // extern "C"
