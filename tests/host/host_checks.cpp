// Host-side checks of the PRODUCT headers that are host-compilable (kd_select.h, jacobi3.h) and a
// lane-by-lane model of the cooperative Hoare-partition formulation used by kd_build.cuh.
// Built by __graft_entry__.build() into tests/host/libhost_checks.so; driven by tests/test_host_logic.py.
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

} // extern "C"
