// TEST INFRASTRUCTURE ONLY — see oracle.h for the contract and the pinning status.
//
// CPU restatement of the reference hot path. Each function cites the reference file:line it
// follows. Compiled with the reference's Release flags (-O3 -DNDEBUG, no -march, no fast-math) so
// float arithmetic is plain IEEE mul/add without FMA contraction, like the reference build.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <unordered_map>
#include <vector>

namespace
{
constexpr std::uint32_t kUnknown = 0U;  // SegmentationLabel::UNKNOWN  (segmentation.hpp:41-46)
constexpr std::uint32_t kGround = 1U;   // SegmentationLabel::GROUND
constexpr std::uint32_t kObstacle = 2U; // SegmentationLabel::OBSTACLE

// ---------------------------------------------------------------------------------------------
// Eigen 3.4 JacobiSVD<Matrix3f>, square real case, restated from the published algorithm
// (Eigen/src/SVD/JacobiSVD.h compute(); Eigen/src/misc/RealSvd2x2.h real_2x2_jacobi_svd();
//  Eigen/src/Jacobi/Jacobi.h makeJacobi(), operator*, transpose(), apply_rotation_in_the_plane()).
// Call site: reference src/segmentation.cpp:87-94, member declared at src/segmentation.hpp:110.
// ---------------------------------------------------------------------------------------------
struct Rot
{
    float c{1.0F};
    float s{0.0F};
};

inline Rot rot_mul(const Rot &a, const Rot &b) // JacobiRotation::operator*
{
    Rot r;
    r.c = a.c * b.c - a.s * b.s;
    r.s = a.c * b.s + a.s * b.c;
    return r;
}

inline Rot rot_transpose(const Rot &a)
{
    Rot r;
    r.c = a.c;
    r.s = -a.s;
    return r;
}

// apply_rotation_in_the_plane on two strided 3-vectors (scalar path: size 3 is not a packet multiple)
inline void rot_apply(float *x, int incx, float *y, int incy, int n, const Rot &j)
{
    if (j.c == 1.0F && j.s == 0.0F)
        return;
    for (int i = 0; i < n; ++i)
    {
        const float xi = x[i * incx];
        const float yi = y[i * incy];
        x[i * incx] = j.c * xi + j.s * yi;
        y[i * incy] = -j.s * xi + j.c * yi;
    }
}

inline Rot make_jacobi(float x, float y, float z) // JacobiRotation::makeJacobi(x, y, z)
{
    Rot r;
    const float deno = 2.0F * std::fabs(y);
    if (deno < std::numeric_limits<float>::min())
    {
        r.c = 1.0F;
        r.s = 0.0F;
        return r;
    }
    const float tau = (x - z) / deno;
    const float w = std::sqrt(tau * tau + 1.0F);
    float t;
    if (tau > 0.0F)
        t = 1.0F / (tau + w);
    else
        t = 1.0F / (tau - w);
    const float sign_t = t > 0.0F ? 1.0F : -1.0F;
    const float n = 1.0F / std::sqrt(t * t + 1.0F);
    r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
    r.c = n;
    return r;
}

// W, V row-major 3x3. Returns number of sweeps, -1 on non-finite input (info() != Success).
int jacobi_svd3(const float a_in[9], float v_out[9], float sv_out[3])
{
    const float precision = 2.0F * std::numeric_limits<float>::epsilon();
    const float consider_as_zero = std::numeric_limits<float>::min();

    float scale = 0.0F;
    for (int i = 0; i < 9; ++i)
    {
        const float v = std::fabs(a_in[i]);
        if (!(v <= scale)) // propagates NaN like maxCoeff<PropagateNaN>
            scale = v;
    }
    if (!std::isfinite(scale))
        return -1;
    if (scale == 0.0F)
        scale = 1.0F;

    float w[9];
    float v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i)
        w[i] = a_in[i] / scale;

    float max_diag = std::max(std::fabs(w[0]), std::max(std::fabs(w[4]), std::fabs(w[8])));

    int sweeps = 0;
    bool finished = false;
    while (!finished)
    {
        finished = true;
        ++sweeps;
        for (int p = 1; p < 3; ++p)
        {
            for (int q = 0; q < p; ++q)
            {
                const float threshold = std::max(consider_as_zero, precision * max_diag);
                if (std::fabs(w[p * 3 + q]) > threshold || std::fabs(w[q * 3 + p]) > threshold)
                {
                    finished = false;
                    // real_2x2_jacobi_svd
                    float m[4] = {w[p * 3 + p], w[p * 3 + q], w[q * 3 + p], w[q * 3 + q]};
                    Rot rot1;
                    const float t = m[0] + m[3];
                    const float d = m[2] - m[1];
                    if (std::fabs(d) < std::numeric_limits<float>::min())
                    {
                        rot1.s = 0.0F;
                        rot1.c = 1.0F;
                    }
                    else
                    {
                        const float u = t / d;
                        const float tmp = std::sqrt(1.0F + u * u);
                        rot1.s = 1.0F / tmp;
                        rot1.c = u / tmp;
                    }
                    rot_apply(&m[0], 1, &m[2], 1, 2, rot1); // m.applyOnTheLeft(0,1,rot1)
                    const Rot j_right = make_jacobi(m[0], m[1], m[3]);
                    const Rot j_left = rot_mul(rot1, rot_transpose(j_right));

                    rot_apply(&w[p * 3], 1, &w[q * 3], 1, 3, j_left);           // applyOnTheLeft(p,q,j_left)
                    rot_apply(&w[p], 3, &w[q], 3, 3, rot_transpose(j_right));   // applyOnTheRight(p,q,j_right)
                    rot_apply(&v[p], 3, &v[q], 3, 3, rot_transpose(j_right));   // V.applyOnTheRight(p,q,j_right)

                    max_diag = std::max(max_diag, std::max(std::fabs(w[p * 3 + p]), std::fabs(w[q * 3 + q])));
                }
            }
        }
        if (sweeps > 1000)
            break; // never observed; guards the test harness against a hang
    }

    float sv[3];
    for (int i = 0; i < 3; ++i)
        sv[i] = std::fabs(w[i * 3 + i]) * scale; // V columns are not sign-flipped (only U's would be)

    for (int i = 0; i < 3; ++i)
    {
        int pos = 0;
        float best = sv[i];
        for (int k = i + 1; k < 3; ++k)
            if (sv[k] > best)
            {
                best = sv[k];
                pos = k - i;
            }
        if (best == 0.0F)
            break;
        if (pos)
        {
            pos += i;
            std::swap(sv[i], sv[pos]);
            for (int r = 0; r < 3; ++r)
                std::swap(v[r * 3 + i], v[r * 3 + pos]);
        }
    }
    std::memcpy(v_out, v, sizeof(v));
    std::memcpy(sv_out, sv, sizeof(sv));
    return sweeps;
}

struct Plane
{
    float a{0}, b{0}, c{0}, d{0};
};

// Segmenter::estimate_plane_coefficients (segmentation.cpp:62-102); xyz is n x 3 row-major.
bool estimate_plane(const std::vector<float> &xyz, std::size_t n, Plane &plane)
{
    if (n < 3)
        return false;
    float cx = 0.0F, cy = 0.0F, cz = 0.0F; // colwise().mean(): sequential column sums / n
    for (std::size_t i = 0; i < n; ++i)
    {
        cx += xyz[i * 3 + 0];
        cy += xyz[i * 3 + 1];
        cz += xyz[i * 3 + 2];
    }
    const float fn = static_cast<float>(n);
    cx /= fn;
    cy /= fn;
    cz /= fn;

    float cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; // centeredᵀ·centered (summation order of Eigen's GEMM unknowable)
    for (std::size_t i = 0; i < n; ++i)
    {
        const float dx = xyz[i * 3 + 0] - cx;
        const float dy = xyz[i * 3 + 1] - cy;
        const float dz = xyz[i * 3 + 2] - cz;
        cov[0] += dx * dx;
        cov[1] += dx * dy;
        cov[2] += dx * dz;
        cov[4] += dy * dy;
        cov[5] += dy * dz;
        cov[8] += dz * dz;
    }
    cov[3] = cov[1];
    cov[6] = cov[2];
    cov[7] = cov[5];
    const float denom = static_cast<float>(n - 1);
    for (float &c : cov)
        c /= denom;

    float v[9], sv[3];
    if (jacobi_svd3(cov, v, sv) < 0)
        return false;
    plane.a = v[0 * 3 + 2];
    plane.b = v[1 * 3 + 2];
    plane.c = v[2 * 3 + 2];
    plane.d = plane.a * cx + plane.b * cy + plane.c * cz;
    return true;
}

struct SegPoint
{
    float x, y, z;
    std::uint32_t index;
};

// Segmenter::extract_initial_seeds (segmentation.cpp:151-217): returns seed local indices in z-sorted order.
void extract_seeds(const std::vector<SegPoint> &seg, const oracle_seg_cfg &cfg, std::vector<std::uint32_t> &ground)
{
    ground.clear();
    if (seg.empty())
        return;
    std::vector<std::uint32_t> order(seg.size());
    std::iota(order.begin(), order.end(), 0U);
    std::sort(order.begin(), order.end(),
              [&seg](std::uint32_t a, std::uint32_t b) -> bool { return seg[a].z < seg[b].z; });

    const float z_min = -1.5F * cfg.sensor_height_m;
    std::size_t z_min_cut = 0;
    for (std::size_t i = 0; i < order.size(); ++i)
        if (seg[order[i]].z > z_min)
        {
            z_min_cut = i;
            break;
        }
    order.erase(order.begin(), order.begin() + static_cast<std::ptrdiff_t>(z_min_cut));
    if (order.empty())
        return;

    float z_mean = 0.0F;
    const std::size_t n_lpr =
        std::min(order.size(), static_cast<std::size_t>(cfg.number_of_lower_point_representatives));
    for (std::size_t i = 0; i < n_lpr; ++i)
        z_mean += seg[order[i]].z;
    z_mean /= n_lpr; // size_t -> float conversion, as in the reference

    const float z_max = z_mean + cfg.initial_seed_threshold;
    std::size_t z_max_cut = 0;
    for (std::size_t i = 0; i < order.size(); ++i)
        if (seg[order[i]].z > z_max)
        {
            z_max_cut = i;
            break;
        }
    ground.assign(order.begin(), order.begin() + static_cast<std::ptrdiff_t>(z_max_cut));
}

// Segmenter::fit_ground_plane (segmentation.cpp:219-309). status: 0 ok, 1 "<3 points", 2 failed.
int fit_ground_plane(const std::vector<SegPoint> &seg, const oracle_seg_cfg &cfg, std::vector<std::uint32_t> &ground,
                     std::vector<std::uint32_t> &obstacle, float *planes /* iters*4 or null */)
{
    ground.clear();
    obstacle.clear();
    const std::uint32_t n = static_cast<std::uint32_t>(seg.size());
    if (n < 3)
        return 1;

    extract_seeds(seg, cfg, ground);

    std::vector<float> ground_xyz;
    for (std::uint32_t it = 0; it < cfg.number_of_iterations; ++it)
    {
        const std::size_t ng = ground.size();
        bool ok = ng >= 3;
        Plane plane;
        if (ok)
        {
            ground_xyz.resize(ng * 3);
            for (std::size_t i = 0; i < ng; ++i)
            {
                const SegPoint &p = seg[ground[i]];
                ground_xyz[i * 3 + 0] = p.x;
                ground_xyz[i * 3 + 1] = p.y;
                ground_xyz[i * 3 + 2] = p.z;
            }
            ok = estimate_plane(ground_xyz, ng, plane);
        }
        if (!ok)
        {
            ground.clear();
            obstacle.resize(n);
            std::iota(obstacle.begin(), obstacle.end(), 0U);
            return 2;
        }
        if (planes)
        {
            planes[it * 4 + 0] = plane.a;
            planes[it * 4 + 1] = plane.b;
            planes[it * 4 + 2] = plane.c;
            planes[it * 4 + 3] = plane.d;
        }
        // distances = P·n − d (row-major GEMV, 3 columns: ((x·a + y·b) + z·c)), then threshold 0.3·‖n‖
        const float norm = std::sqrt((plane.a * plane.a + plane.b * plane.b) + plane.c * plane.c);
        const float thr = cfg.orthogonal_distance_threshold * norm;
        ground.clear();
        obstacle.clear();
        for (std::uint32_t i = 0; i < n; ++i)
        {
            const float dist = ((seg[i].x * plane.a + seg[i].y * plane.b) + seg[i].z * plane.c) - plane.d;
            if (dist < thr)
                ground.push_back(i);
            else
                obstacle.push_back(i);
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// k-d tree (kdtree.hpp) restated over an implicit tree: node of range [b,e) sits at b+(e-b)/2.
// ---------------------------------------------------------------------------------------------
struct KdNode
{
    float p[3];
    std::uint32_t index;
};

inline float dist_sqr(const float *a, const float *b) // kdtree.hpp:145-163 (recursive template, right fold)
{
    const float d0 = (a[0] - b[0]) * (a[0] - b[0]);
    const float d1 = (a[1] - b[1]) * (a[1] - b[1]);
    const float d2 = (a[2] - b[2]) * (a[2] - b[2]);
    return d0 + (d1 + (d2 + 0.0F));
}

// ---- hand transcription of libstdc++ 13 std::nth_element (what the device must emulate) ----
namespace tx
{
struct Less
{
    int axis;
    bool operator()(const KdNode &a, const KdNode &b) const { return a.p[axis] < b.p[axis]; }
};

inline void push_heap_(KdNode *first, long hole, long top, KdNode value, const Less &comp)
{
    long parent = (hole - 1) / 2;
    while (hole > top && comp(first[parent], value))
    {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

inline void adjust_heap_(KdNode *first, long hole, long len, KdNode value, const Less &comp)
{
    const long top = hole;
    long second = hole;
    while (second < (len - 1) / 2)
    {
        second = 2 * (second + 1);
        if (comp(first[second], first[second - 1]))
            --second;
        first[hole] = first[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2)
    {
        second = 2 * (second + 1);
        first[hole] = first[second - 1];
        hole = second - 1;
    }
    push_heap_(first, hole, top, value, comp);
}

inline void make_heap_(KdNode *first, KdNode *last, const Less &comp)
{
    if (last - first < 2)
        return;
    const long len = last - first;
    long parent = (len - 2) / 2;
    while (true)
    {
        KdNode value = first[parent];
        adjust_heap_(first, parent, len, value, comp);
        if (parent == 0)
            return;
        --parent;
    }
}

inline void heap_select_(KdNode *first, KdNode *middle, KdNode *last, const Less &comp)
{
    make_heap_(first, middle, comp);
    for (KdNode *i = middle; i < last; ++i)
        if (comp(*i, *first))
        {
            KdNode value = *i; // __pop_heap(first, middle, i)
            *i = *first;
            adjust_heap_(first, 0, middle - first, value, comp);
        }
}

inline void insertion_sort_(KdNode *first, KdNode *last, const Less &comp)
{
    if (first == last)
        return;
    for (KdNode *i = first + 1; i != last; ++i)
    {
        if (comp(*i, *first))
        {
            KdNode val = *i;
            for (KdNode *k = i; k != first; --k)
                *k = *(k - 1);
            *first = val;
        }
        else
        {
            KdNode val = *i;
            KdNode *lastp = i;
            KdNode *next = i - 1;
            while (comp(val, *next))
            {
                *lastp = *next;
                lastp = next;
                --next;
            }
            *lastp = val;
        }
    }
}

inline void move_median_to_first_(KdNode *result, KdNode *a, KdNode *b, KdNode *c, const Less &comp)
{
    if (comp(*a, *b))
    {
        if (comp(*b, *c))
            std::swap(*result, *b);
        else if (comp(*a, *c))
            std::swap(*result, *c);
        else
            std::swap(*result, *a);
    }
    else if (comp(*a, *c))
        std::swap(*result, *a);
    else if (comp(*b, *c))
        std::swap(*result, *c);
    else
        std::swap(*result, *b);
}

inline KdNode *unguarded_partition_(KdNode *first, KdNode *last, KdNode *pivot, const Less &comp)
{
    while (true)
    {
        while (comp(*first, *pivot))
            ++first;
        --last;
        while (comp(*pivot, *last))
            --last;
        if (!(first < last))
            return first;
        std::swap(*first, *last);
        ++first;
    }
}

inline void nth_element_(KdNode *first, KdNode *nth, KdNode *last, const Less &comp)
{
    if (first == last || nth == last)
        return;
    long n = last - first;
    long depth_limit = 0;
    while (n > 1) // std::__lg
    {
        n >>= 1;
        ++depth_limit;
    }
    depth_limit *= 2;
    while (last - first > 3)
    {
        if (depth_limit == 0)
        {
            heap_select_(first, nth + 1, last, comp);
            std::swap(*first, *nth);
            return;
        }
        --depth_limit;
        KdNode *mid = first + (last - first) / 2;
        move_median_to_first_(first, first + 1, mid, last - 1, comp);
        KdNode *cut = unguarded_partition_(first + 1, last, first, comp);
        if (cut <= nth)
            first = cut;
        else
            last = cut;
    }
    insertion_sort_(first, last, comp);
}
} // namespace tx

void kd_build(std::vector<KdNode> &nodes, int mode) // KDTree::rebuild (kdtree.hpp:174-225)
{
    struct Range
    {
        std::size_t b, e;
        std::uint32_t depth;
    };
    std::vector<Range> stack;
    stack.push_back({0, nodes.size(), 0U});
    while (!stack.empty())
    {
        const Range r = stack.back();
        stack.pop_back();
        if (r.b >= r.e)
            continue;
        const int axis = static_cast<int>(r.depth % 3U);
        const std::size_t mid = r.b + (r.e - r.b) / 2;
        if (mode == 0)
            std::nth_element(nodes.begin() + static_cast<std::ptrdiff_t>(r.b),
                             nodes.begin() + static_cast<std::ptrdiff_t>(mid),
                             nodes.begin() + static_cast<std::ptrdiff_t>(r.e),
                             [axis](const KdNode &a, const KdNode &b) -> bool { return a.p[axis] < b.p[axis]; });
        else
            tx::nth_element_(nodes.data() + r.b, nodes.data() + mid, nodes.data() + r.e, tx::Less{axis});
        if (mid > r.b)
            stack.push_back({r.b, mid, r.depth + 1});
        if (mid + 1 < r.e)
            stack.push_back({mid + 1, r.e, r.depth + 1});
    }
}

// pre-order of the implicit tree: node, left subtree, right subtree (radius_search pushes right
// then left on a LIFO stack, kdtree.hpp:324-333)
void kd_preorder(const std::vector<KdNode> &nodes, std::uint32_t *order_out)
{
    struct Range
    {
        std::size_t b, e;
    };
    std::vector<Range> stack;
    stack.push_back({0, nodes.size()});
    std::size_t rank = 0;
    while (!stack.empty())
    {
        const Range r = stack.back();
        stack.pop_back();
        if (r.b >= r.e)
            continue;
        const std::size_t mid = r.b + (r.e - r.b) / 2;
        order_out[rank++] = nodes[mid].index;
        stack.push_back({mid + 1, r.e});
        stack.push_back({r.b, mid});
    }
}

// KDTree::radius_search (kdtree.hpp:292-341), sort_ == false
void kd_radius_search(const std::vector<KdNode> &nodes, const float *target, float prox_sqr,
                      std::vector<std::pair<std::uint32_t, float>> &neigh)
{
    struct Item
    {
        std::size_t b, e;
        std::uint32_t axis;
    };
    static thread_local std::vector<Item> stack;
    neigh.clear();
    stack.clear();
    stack.push_back({0, nodes.size(), 0U});
    while (!stack.empty())
    {
        const Item it = stack.back();
        stack.pop_back();
        if (it.b >= it.e)
            continue; // nullptr child
        const std::size_t mid = it.b + (it.e - it.b) / 2;
        const KdNode &node = nodes[mid];
        const float dist = dist_sqr(target, node.p);
        if (dist <= prox_sqr)
            neigh.emplace_back(node.index, dist);
        const std::uint32_t next_axis = (it.axis + 1U) % 3U;
        const float delta = node.p[it.axis] - target[it.axis];
        const float abs_delta_sqr = delta * delta;
        if (abs_delta_sqr <= prox_sqr)
        {
            stack.push_back({mid + 1, it.e, next_axis});
            stack.push_back({it.b, mid, next_axis});
        }
        else if (delta > 0)
            stack.push_back({it.b, mid, next_axis});
        else
            stack.push_back({mid + 1, it.e, next_axis});
    }
}

std::vector<KdNode> load_nodes(const float *pts, std::uint32_t m, std::uint32_t stride)
{
    std::vector<KdNode> nodes(m);
    for (std::uint32_t i = 0; i < m; ++i)
    {
        const float *p = pts + static_cast<std::size_t>(i) * stride;
        nodes[i] = KdNode{{p[0], p[1], p[2]}, i};
    }
    return nodes;
}

// ---------------------------------------------------------------------------------------------
// voxel grid used by the CC pass and by the device-formulation model
// ---------------------------------------------------------------------------------------------
struct Grid
{
    double inv_cell;
    double origin[3];
    std::vector<std::uint32_t> order;      // point indices sorted by cell key
    std::vector<std::uint64_t> key_of;     // per point
    std::unordered_map<std::uint64_t, std::pair<std::uint32_t, std::uint32_t>> cells; // key -> [start,end) in order

    static std::uint64_t pack(std::int64_t cx, std::int64_t cy, std::int64_t cz)
    {
        return (static_cast<std::uint64_t>(cz) << 42) | (static_cast<std::uint64_t>(cy) << 21) |
               static_cast<std::uint64_t>(cx);
    }
    void cell_of(const float *p, std::int64_t c[3]) const
    {
        for (int a = 0; a < 3; ++a)
            c[a] = static_cast<std::int64_t>(std::floor((static_cast<double>(p[a]) - origin[a]) * inv_cell)) + 1;
    }
    void build(const float *pts, std::uint32_t m, std::uint32_t stride, float dsq)
    {
        const double cell = std::sqrt(static_cast<double>(dsq)) * 1.001;
        inv_cell = 1.0 / cell;
        for (int a = 0; a < 3; ++a)
            origin[a] = std::numeric_limits<double>::infinity();
        for (std::uint32_t i = 0; i < m; ++i)
            for (int a = 0; a < 3; ++a)
                origin[a] = std::min(origin[a], static_cast<double>(pts[static_cast<std::size_t>(i) * stride + a]));
        key_of.resize(m);
        order.resize(m);
        for (std::uint32_t i = 0; i < m; ++i)
        {
            std::int64_t c[3];
            cell_of(pts + static_cast<std::size_t>(i) * stride, c);
            key_of[i] = pack(c[0], c[1], c[2]);
            order[i] = i;
        }
        std::stable_sort(order.begin(), order.end(),
                         [this](std::uint32_t a, std::uint32_t b) { return key_of[a] < key_of[b]; });
        cells.clear();
        cells.reserve(m);
        std::uint32_t s = 0;
        while (s < m)
        {
            std::uint32_t e = s + 1;
            while (e < m && key_of[order[e]] == key_of[order[s]])
                ++e;
            cells.emplace(key_of[order[s]], std::make_pair(s, e));
            s = e;
        }
    }
    template <typename F> void for_each_candidate(const float *p, F &&f) const
    {
        std::int64_t c[3];
        cell_of(p, c);
        for (std::int64_t dz = -1; dz <= 1; ++dz)
            for (std::int64_t dy = -1; dy <= 1; ++dy)
                for (std::int64_t dx = -1; dx <= 1; ++dx)
                {
                    const auto it = cells.find(pack(c[0] + dx, c[1] + dy, c[2] + dz));
                    if (it == cells.end())
                        continue;
                    for (std::uint32_t s = it->second.first; s < it->second.second; ++s)
                        f(order[s]);
                }
    }
};

std::uint32_t uf_find(std::vector<std::uint32_t> &parent, std::uint32_t x)
{
    while (parent[x] != x)
    {
        parent[x] = parent[parent[x]];
        x = parent[x];
    }
    return x;
}
} // namespace

extern "C"
{

void oracle_seg_cfg_default(oracle_seg_cfg *cfg)
{
    cfg->sensor_height_m = 1.73F;
    cfg->orthogonal_distance_threshold = 0.3F;
    cfg->initial_seed_threshold = 0.6F;
    cfg->number_of_iterations = 3U;
    cfg->number_of_planar_partitions = 2U;
    cfg->number_of_lower_point_representatives = 5000U;
}

void oracle_clu_cfg_default(oracle_clu_cfg *cfg)
{
    cfg->distance_squared = 0.18F;
    cfg->cluster_quality = 0.5F;
    cfg->min_cluster_size = 4U;
    cfg->max_cluster_size = std::numeric_limits<std::uint32_t>::max();
}

int oracle_jacobi_svd3(const float *a, float *v_out, float *sv_out)
{
    return jacobi_svd3(a, v_out, sv_out);
}

int oracle_segment(const float *pts, std::uint32_t n, std::uint32_t stride, const oracle_seg_cfg *cfg, int tie_mode,
                   std::uint32_t *labels, std::uint32_t *ground_idx, std::uint32_t *n_ground,
                   std::uint32_t *obstacle_idx, std::uint32_t *n_obstacle, float *planes_out,
                   std::int32_t *seg_status_out)
{
    *n_ground = 0;
    *n_obstacle = 0;
    const std::uint32_t P = cfg->number_of_planar_partitions;
    if (P == 0U)
        return -1; // the reference divides by zero here (segmentation.cpp:124)
    if (planes_out)
        for (std::size_t i = 0; i < static_cast<std::size_t>(P) * cfg->number_of_iterations * 4; ++i)
            planes_out[i] = std::numeric_limits<float>::quiet_NaN();
    if (seg_status_out)
        for (std::uint32_t s = 0; s < P; ++s)
            seg_status_out[s] = 1;
    if (n == 0U)
        return 0; // segmentation.cpp:319-323

    // form_planar_partitions (segmentation.cpp:104-149)
    std::vector<std::uint32_t> sorted(n);
    std::iota(sorted.begin(), sorted.end(), 0U);
    auto less_x = [pts, stride](std::uint32_t a, std::uint32_t b) -> bool {
        return pts[static_cast<std::size_t>(a) * stride] < pts[static_cast<std::size_t>(b) * stride];
    };
    if (tie_mode == 0)
        std::sort(sorted.begin(), sorted.end(), less_x);
    else
        std::stable_sort(sorted.begin(), sorted.end(), less_x);

    const std::size_t per = n / P;
    std::size_t lo = 0, hi = per;
    std::vector<SegPoint> seg;
    std::vector<std::uint32_t> ground, obstacle;
    for (std::uint32_t s = 0; s < P; ++s)
    {
        seg.clear();
        for (std::size_t i = lo; i < hi; ++i)
        {
            const std::uint32_t pi = sorted[i];
            const float *p = pts + static_cast<std::size_t>(pi) * stride;
            seg.push_back(SegPoint{p[0], p[1], p[2], pi});
        }
        lo = hi;
        hi = std::min(lo + per, static_cast<std::size_t>(n));

        const int status = fit_ground_plane(
            seg, *cfg, ground, obstacle,
            planes_out ? planes_out + static_cast<std::size_t>(s) * cfg->number_of_iterations * 4 : nullptr);
        if (seg_status_out)
            seg_status_out[s] = status;
        for (const std::uint32_t g : ground) // segmentation.cpp:331-336
        {
            labels[seg[g].index] = kGround;
            ground_idx[(*n_ground)++] = seg[g].index;
        }
        for (const std::uint32_t o : obstacle) // segmentation.cpp:338-343
        {
            labels[seg[o].index] = kObstacle;
            obstacle_idx[(*n_obstacle)++] = seg[o].index;
        }
    }
    (void)kUnknown;
    return 0;
}

int oracle_kd_build(const float *pts, std::uint32_t m, std::uint32_t stride, int mode, std::uint32_t *slot_out)
{
    std::vector<KdNode> nodes = load_nodes(pts, m, stride);
    if (m)
        kd_build(nodes, mode);
    for (std::uint32_t i = 0; i < m; ++i)
        slot_out[i] = nodes[i].index;
    return 0;
}

int oracle_kd_order(const float *pts, std::uint32_t m, std::uint32_t stride, int mode, std::uint32_t *order_out)
{
    std::vector<KdNode> nodes = load_nodes(pts, m, stride);
    if (m)
    {
        kd_build(nodes, mode);
        kd_preorder(nodes, order_out);
    }
    return 0;
}

int oracle_cluster(const float *pts, std::uint32_t m, std::uint32_t stride, const oracle_clu_cfg *cfg,
                   std::int32_t *labels)
{
    constexpr std::int32_t kUndefined = std::numeric_limits<std::int32_t>::lowest();
    for (std::uint32_t i = 0; i < m; ++i)
        labels[i] = kUndefined; // clustering.cpp:50
    if (m == 0U)
        return 0;

    std::vector<KdNode> nodes = load_nodes(pts, m, stride);
    kd_build(nodes, 0);
    std::vector<char> removed(m, 0);
    // clustering.cpp:66-67 — evaluated in double
    const double inner = std::pow(1.0 - cfg->cluster_quality, 2) * cfg->distance_squared;

    std::vector<std::pair<std::uint32_t, float>> neigh;
    std::vector<std::uint32_t> queue; // FIFO: head index into a growing vector
    std::vector<std::uint32_t> indices;
    std::int32_t label = 0;
    for (std::uint32_t i = 0; i < m; ++i)
    {
        if (removed[i])
            continue;
        queue.clear();
        std::size_t head = 0;
        queue.push_back(i);
        indices.clear();
        while (head < queue.size())
        {
            const std::uint32_t j = queue[head++];
            if (removed[j])
                continue;
            kd_radius_search(nodes, pts + static_cast<std::size_t>(j) * stride, cfg->distance_squared, neigh);
            for (const auto &[k, dist] : neigh)
            {
                if (removed[k])
                    continue;
                labels[k] = label;
                indices.push_back(k);
                if (dist <= inner)
                    removed[k] = 1;
                else
                    queue.push_back(k);
            }
        }
        if (indices.size() < cfg->min_cluster_size || indices.size() > cfg->max_cluster_size)
            for (const std::uint32_t k : indices)
                labels[k] = -1;
        else
            ++label;
    }
    return 0;
}

int oracle_cc(const float *pts, std::uint32_t m, std::uint32_t stride, float dsq, std::uint32_t *root_out)
{
    if (m == 0U)
        return 0;
    Grid grid;
    grid.build(pts, m, stride, dsq);
    std::vector<std::uint32_t> parent(m);
    std::iota(parent.begin(), parent.end(), 0U);
    for (std::uint32_t i = 0; i < m; ++i)
    {
        const float *p = pts + static_cast<std::size_t>(i) * stride;
        grid.for_each_candidate(p, [&](std::uint32_t k) {
            if (k <= i)
                return;
            if (dist_sqr(p, pts + static_cast<std::size_t>(k) * stride) <= dsq)
            {
                std::uint32_t a = uf_find(parent, i), b = uf_find(parent, k);
                if (a != b)
                {
                    if (a < b)
                        parent[b] = a;
                    else
                        parent[a] = b;
                }
            }
        });
    }
    for (std::uint32_t i = 0; i < m; ++i)
        root_out[i] = uf_find(parent, i);
    return 0;
}

int oracle_cluster_model(const float *pts, std::uint32_t m, std::uint32_t stride, const oracle_clu_cfg *cfg,
                         const std::uint32_t *rank_of, std::int32_t *labels, std::uint64_t *stats)
{
    std::uint64_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (m == 0U)
    {
        if (stats)
            std::memcpy(stats, st, sizeof(st));
        return 0;
    }
    const float dsq = cfg->distance_squared;
    const double inner = std::pow(1.0 - cfg->cluster_quality, 2) * cfg->distance_squared;

    Grid grid;
    grid.build(pts, m, stride, dsq);
    std::vector<std::uint32_t> root(m);
    oracle_cc(pts, m, stride, dsq, root.data());

    // members of each r-component in ascending index order
    std::vector<std::uint32_t> comp_members(m);
    std::iota(comp_members.begin(), comp_members.end(), 0U);
    std::stable_sort(comp_members.begin(), comp_members.end(),
                     [&root](std::uint32_t a, std::uint32_t b) { return root[a] < root[b]; });

    std::vector<char> removed(m, 0), queued(m, 0);
    std::vector<std::uint32_t> seed_of(m, 0xFFFFFFFFU);
    std::vector<char> seed_valid(m, 0);
    std::vector<std::uint32_t> queue;
    struct Hit
    {
        std::uint32_t rank, idx;
    };
    std::vector<Hit> pushes;

    std::size_t s = 0;
    while (s < m)
    {
        std::size_t e = s + 1;
        while (e < m && root[comp_members[e]] == root[comp_members[s]])
            ++e;
        ++st[4];
        st[5] = std::max<std::uint64_t>(st[5], e - s);
        std::uint64_t comp_expansions = 0;
        for (std::size_t t = s; t < e; ++t) // replay of clustering.cpp:69-124 restricted to one component
        {
            const std::uint32_t i = comp_members[t];
            if (removed[i])
                continue;
            ++st[0];
            queue.clear();
            std::size_t head = 0;
            queue.push_back(i);
            queued[i] = 1;
            std::uint64_t touched = 0; // indices_.size(): pushes WITH multiplicity
            while (head < queue.size())
            {
                st[3] = std::max<std::uint64_t>(st[3], queue.size() - head);
                const std::uint32_t j = queue[head++];
                if (removed[j])
                    continue;
                ++st[1];
                ++comp_expansions;
                const float *pj = pts + static_cast<std::size_t>(j) * stride;
                pushes.clear();
                grid.for_each_candidate(pj, [&](std::uint32_t k) {
                    const float d = dist_sqr(pj, pts + static_cast<std::size_t>(k) * stride);
                    if (!(d <= dsq))
                        return;
                    ++st[7];
                    if (removed[k])
                        return;
                    seed_of[k] = i;
                    ++touched;
                    if (static_cast<double>(d) <= inner)
                        removed[k] = 1;
                    else if (!queued[k]) // duplicates in the reference FIFO are no-ops when popped
                    {
                        queued[k] = 1;
                        pushes.push_back(Hit{rank_of[k], k});
                    }
                });
                std::sort(pushes.begin(), pushes.end(), [](const Hit &a, const Hit &b) { return a.rank < b.rank; });
                for (const Hit &h : pushes)
                    queue.push_back(h.idx);
                st[2] += pushes.size();
            }
            seed_valid[i] = !(touched < cfg->min_cluster_size || touched > cfg->max_cluster_size);
        }
        st[6] = std::max(st[6], comp_expansions);
        s = e;
    }

    // label compaction: label = number of valid seeds with a smaller index
    std::vector<std::int32_t> seed_label(m, -1);
    std::int32_t next = 0;
    for (std::uint32_t i = 0; i < m; ++i)
        if (seed_of[i] == i && seed_valid[i])
            seed_label[i] = next++;
    for (std::uint32_t i = 0; i < m; ++i)
        labels[i] = seed_of[i] == 0xFFFFFFFFU ? std::numeric_limits<std::int32_t>::lowest() : seed_label[seed_of[i]];
    if (stats)
        std::memcpy(stats, st, sizeof(st));
    return 0;
}

std::uint32_t oracle_canonicalise(const std::int32_t *labels, std::uint32_t m, std::int32_t *out)
{
    std::unordered_map<std::int32_t, std::int32_t> remap;
    std::int32_t next = 0;
    for (std::uint32_t i = 0; i < m; ++i)
    {
        if (labels[i] < 0)
        {
            out[i] = labels[i];
            continue;
        }
        auto it = remap.find(labels[i]);
        if (it == remap.end())
            it = remap.emplace(labels[i], next++).first;
        out[i] = it->second;
    }
    return static_cast<std::uint32_t>(next);
}

std::uint64_t oracle_fnv1a64(const void *data, std::uint64_t nbytes)
{
    const unsigned char *p = static_cast<const unsigned char *>(data);
    std::uint64_t h = 1469598103934665603ULL;
    for (std::uint64_t i = 0; i < nbytes; ++i)
    {
        h ^= p[i];
        h *= 1099511628211ULL;
    }
    return h;
}

} // extern "C"
