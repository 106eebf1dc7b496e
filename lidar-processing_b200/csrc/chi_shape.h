// Concave outline of one cluster (the part for clusters of 20 points and more), host- and
// device-compilable: the sequential core that chi_shape.cuh runs per cluster and tests/host/host_checks.cpp runs on
// the CPU against the unmodified reference.
//
// What it re-enacts, value for value in float64 (no FMA contraction; IEEE +, -, *, /, sqrt):
//   * geometry::ConcaveHull<float>::constructConcaveHull with chi = 0.2, as findOrderedConcaveOutlines calls it
//     (reference src/polygon_simplification.cpp:119-140, Concave-Hull/concave_hull.hpp:96-193): a max-heap of the
//     boundary edges by length (std::push_heap / std::pop_heap of libstdc++ 13: the order among edges of EQUAL
//     length decides which one is eroded first, and mm-quantised clusters are full of equal lengths), erosion of
//     every boundary edge longer than chi * max + (1 - chi) * min whose opposite vertex is not on the boundary yet,
//     then the walk along hull_next from hull_start with the start repeated at the end (delaunator.cpp:693-707);
//   * delaunator::Delaunator (Concave-Hull/delaunator.cpp:214-487, 520-633): seed point nearest to the bounding-box
//     centre, its nearest distinct neighbour, the third point of the smallest circumcircle, points in the order of
//     their distance from that circumcentre, the advancing hull with its pseudo-angle hash, edge legalisation with an
//     explicit stack. Every predicate keeps the reference's thresholds (epsilon-guarded orientation and in-circle
//     tests, `<` against `<=` where the two circumradius overloads differ).
//   * the order of the points: the reference sorts indices with std::sort by distance only. Where all distances
//     differ, or equal distances belong to coinciding points (value-identical, so their order is invisible in the
//     outline), any sort gives the reference's sequence and the device sorts in parallel by (distance, index). A
//     cluster in which two DIFFERENT points are exactly equally far away takes chi_introsort, a step-by-step
//     re-enactment of libstdc++ 13's std::sort (introsort loop with depth limit 2 * floor(log2 n), median of three
//     moved to the front, unguarded Hoare partition, heap sort when the limit is hit, final insertion sort with its
//     16-element threshold).
//
// Cases in which the reference itself does not deliver: it throws "not triangulation" when every third point is
// collinear with the first two (delaunator.cpp:299; the node goes down with it), and it reads out of bounds when all
// points coincide. Both are reported per cluster (kChiErrCollinear / kChiErrCoincident) with 0 vertices.
#pragma once

#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define LB_CHI_HD __host__ __device__ inline
#else
#define LB_CHI_HD inline
#endif

namespace lb
{

constexpr uint32_t kChiNone = 0xFFFFFFFFu;
constexpr uint32_t kChiOk = 0u, kChiErrCollinear = 1u, kChiErrCoincident = 2u, kChiErrGuard = 3u;
constexpr double kChiEps = DBL_EPSILON;
constexpr double kChiFactor = 0.2; // geometry::ConcaveHull hull(coordinates, 0.2), polygon_simplification.cpp:132

// One record per half-edge: its origin vertex (triangles[e]) WITH that vertex's coordinates, and its twin
// (half_edges[e]). The legalisation reads what it needs about a triangle with three 16-byte loads whose addresses it
// knows up front, and about the triangle across an edge with a fourth - instead of half_edges[] -> triangles[] ->
// coords[] one after the other: on a cluster of 20 000 points every one of those is an L2 round trip of ~290 cycles,
// and the sweep is a chain of them (profiles/r02_micro_l1_read_after_write.txt, r02_chi_outline_task_phases.txt).
struct alignas(16) ChiEdge
{
    uint32_t twin;
    uint32_t v;
    float x, y; // the cluster's float coordinates; the arithmetic widens them like the reference (static_cast<double>)
};

// One record per point: hull_prev / hull_next / hull_tri, its getHashKey and its coordinates (widened once).
struct alignas(16) ChiNode
{
    uint32_t prev, next, tri, key;
    double x, y;
};

constexpr uint32_t kChiLocalStack = 32u; // legalisation stack entries kept in thread-local memory (deeper ones spill)

// Per-cluster working set. Every array lives in one block of chi_layout(n).bytes bytes.
struct ChiWork
{
    ChiNode *node;    // [n]
    uint32_t n;
    double *dist;     // [n + 1] squared distance from the seed circumcentre; later: stack spill, then heap lengths
    uint32_t *ids;    // [n + 1] point order; later: heap edges
    ChiEdge *edge;    // [3 * (2n - 5)] triangles + half_edges
    uint32_t *hash;   // [hash_size] (the device keeps it in shared memory when it fits)
    uint8_t *onb;     // [n] boundary_set of concave_hull.hpp:110
    uint32_t hash_size;
    uint32_t n_half;  // triangles.size()
    bool overflow;    // guard: more triangles than a triangulation can have
    bool tri_moved;   // legalize moved a hull_tri entry (delaunator.cpp:592-606): copies of .tri held in registers are stale
    uint32_t hull_start;
    uint32_t i0, i1, i2;
    double cx, cy;    // m_center
    double span;
    double s0x, s0y, s1x, s1y, s2x, s2y; // seed triangle coordinates (after the orientation swap)
};

struct ChiLayout
{
    size_t dist, ids, edge, node, hash, onb, bytes;
    uint32_t hash_size;
};

LB_CHI_HD size_t chi_align16(size_t v)
{
    return (v + 15u) & ~static_cast<size_t>(15u);
}

// m_hash_size = ceil(sqrt(n)) (delaunator.cpp:344)
LB_CHI_HD uint32_t chi_hash_size(uint32_t n)
{
    return static_cast<uint32_t>(ceil(sqrt(static_cast<double>(n))));
}

// 141 n + 4 sqrt(n) - 138 bytes plus alignment: below 144 n for every n >= 20 (tests/test_host_logic.py)
LB_CHI_HD ChiLayout chi_layout(uint32_t n)
{
    ChiLayout l;
    const size_t np = n, t3 = n >= 3u ? 3u * (2u * np - 5u) : 3u;
    l.hash_size = chi_hash_size(n);
    size_t at = 0;
    l.edge = at;
    at = chi_align16(at + sizeof(ChiEdge) * t3);
    l.node = at;
    at = chi_align16(at + sizeof(ChiNode) * np);
    l.dist = at;
    at = chi_align16(at + 8u * (np + 1u));
    l.ids = at;
    at = chi_align16(at + 4u * (np + 1u));
    l.hash = at;
    at = chi_align16(at + 4u * l.hash_size);
    l.onb = at;
    at = chi_align16(at + np);
    l.bytes = at;
    return l;
}

LB_CHI_HD void chi_bind(ChiWork &w, unsigned char *block, const ChiLayout &l, uint32_t n)
{
    w.n = n;
    w.edge = reinterpret_cast<ChiEdge *>(block + l.edge);
    w.node = reinterpret_cast<ChiNode *>(block + l.node);
    w.dist = reinterpret_cast<double *>(block + l.dist);
    w.ids = reinterpret_cast<uint32_t *>(block + l.ids);
    w.hash = reinterpret_cast<uint32_t *>(block + l.hash);
    w.onb = block + l.onb;
    w.hash_size = l.hash_size;
    w.n_half = 0u;
    w.overflow = false;
    w.tri_moved = false;
}

LB_CHI_HD double chi_px(const ChiWork &w, uint32_t i)
{
    return w.node[i].x;
}

LB_CHI_HD double chi_py(const ChiWork &w, uint32_t i)
{
    return w.node[i].y;
}

// 16-byte record moves as one load / one store
LB_CHI_HD ChiEdge chi_load_edge(const ChiEdge *p)
{
#ifdef __CUDA_ARCH__
    const uint4 r = *reinterpret_cast<const uint4 *>(p);
    ChiEdge e;
    e.twin = r.x;
    e.v = r.y;
    e.x = __uint_as_float(r.z);
    e.y = __uint_as_float(r.w);
    return e;
#else
    return *p;
#endif
}

LB_CHI_HD void chi_store_edge(ChiEdge *p, const ChiEdge &e)
{
#ifdef __CUDA_ARCH__
    *reinterpret_cast<uint4 *>(p) = make_uint4(e.twin, e.v, __float_as_uint(e.x), __float_as_uint(e.y));
#else
    *p = e;
#endif
}

LB_CHI_HD ChiNode chi_load_node(const ChiNode *p)
{
#ifdef __CUDA_ARCH__
    const uint4 a = *reinterpret_cast<const uint4 *>(p);
    const double2 b = *reinterpret_cast<const double2 *>(reinterpret_cast<const unsigned char *>(p) + 16);
    ChiNode nd;
    nd.prev = a.x;
    nd.next = a.y;
    nd.tri = a.z;
    nd.key = a.w;
    nd.x = b.x;
    nd.y = b.y;
    return nd;
#else
    return *p;
#endif
}

// Point::equal(a, b, span) (delaunator.hpp:56-61): squared distance / span < epsilon. The division only decides when
// the squared distance is within a factor of four of epsilon * span; beyond that the quotient is >= 4 eps (1 - 2^-52).
LB_CHI_HD bool chi_near(double d2, double span, double span_4eps)
{
    if (d2 > span_4eps)
        return false;
    return d2 / span < kChiEps;
}

// ---- predicates (delaunator.cpp:40-212) -----------------------------------------------------------------------------

// Point::distanceSquared / distanceSquared
LB_CHI_HD double chi_dist2(double ax, double ay, double bx, double by)
{
    const double dx = ax - bx;
    const double dy = ay - by;
    return dx * dx + dy * dy;
}

// counterclockwise(p, q, r) (delaunator.cpp:122-149): det of (q - p, r - p) beyond +epsilon
LB_CHI_HD bool chi_ccw(double px, double py, double qx, double qy, double rx, double ry)
{
    const double ux = qx - px, uy = qy - py;
    const double vx = rx - px, vy = ry - py;
    const double det = ux * vy - uy * vx;
    return det > kChiEps; // (|det| <= eps -> false is implied)
}

// getCircumRadius(const Point&, const Point&, const Point&) (delaunator.cpp:48-66): used for the third seed point
LB_CHI_HD double chi_circumradius2(double ax, double ay, double bx, double by, double cx, double cy)
{
    const double dx = bx - ax, dy = by - ay;
    const double ex = cx - ax, ey = cy - ay;
    const double det = dx * ey - dy * ex;
    if (fabs(det) < kChiEps)
        return DBL_MAX;
    const double bl = dx * dx + dy * dy;
    const double cl = ex * ex + ey * ey;
    const double rx = ((ey * bl - dy * cl) * 0.5) / det;
    const double ry = ((dx * cl - ex * bl) * 0.5) / det;
    return rx * rx + ry * ry;
}

// getCircumCenter (delaunator.cpp:151-174)
LB_CHI_HD void chi_circumcentre(double ax, double ay, double bx, double by, double cx, double cy, double &ox, double &oy)
{
    const double dx = bx - ax, dy = by - ay;
    const double ex = cx - ax, ey = cy - ay;
    const double d = dx * ey - dy * ex;
    if (fabs(d) <= kChiEps)
    {
        ox = DBL_MAX;
        oy = DBL_MAX;
        return;
    }
    const double bl = dx * dx + dy * dy;
    const double cl = ex * ex + ey * ey;
    ox = ax + ((ey * bl - dy * cl) * 0.5) / d;
    oy = ay + ((dx * cl - ex * bl) * 0.5) / d;
}

// isInsideCircumCircle (delaunator.cpp:176-194)
LB_CHI_HD bool chi_in_circle(double ax, double ay, double bx, double by, double cx, double cy, double px, double py)
{
    const double dx = ax - px, dy = ay - py;
    const double ex = bx - px, ey = by - py;
    const double fx = cx - px, fy = cy - py;
    const double ap = dx * dx + dy * dy;
    const double bp = ex * ex + ey * ey;
    const double cp = fx * fx + fy * fy;
    return (dx * (ey * cp - bp * fy) - dy * (ex * cp - bp * fx) + ap * (ex * fy - ey * fx)) < -kChiEps;
}

// checkPointsEqual (delaunator.cpp:198-201)
LB_CHI_HD bool chi_same(double x1, double y1, double x2, double y2)
{
    return fabs(x1 - x2) <= kChiEps && fabs(y1 - y2) <= kChiEps;
}

// getHashKey (delaunator.cpp:635-642) with pseudoAngle (:204-208) and fastModulus (:17-20)
LB_CHI_HD uint32_t chi_hash_key(const ChiWork &w, double x, double y)
{
    const double dx = x - w.cx;
    const double dy = y - w.cy;
    const double p = dx / (fabs(dx) + fabs(dy));
    const double a = ((dy > 0.0) ? (3.0 - p) : (1.0 + p)) / 4.0;
    const double fl = floor(a * static_cast<double>(w.hash_size));
    // std::llround of a NaN (a point on the circumcentre itself) is what x86-64 makes of it: LLONG_MIN
    const unsigned long long k = (fl != fl) ? 0x8000000000000000ull : static_cast<unsigned long long>(static_cast<long long>(fl));
    return static_cast<uint32_t>(k >= w.hash_size ? k % w.hash_size : k);
}

// ---- libstdc++ 13 heap primitives on (edge, length) pairs, comparator = "length less" (concave_hull.hpp:91-94) ------

// std::__push_heap: the value climbs while its parent is smaller
LB_CHI_HD void chi_heap_sift_up(uint32_t *he, double *hl, uint32_t hole, uint32_t top, uint32_t ve, double vl)
{
    while (hole > top)
    {
        const uint32_t parent = (hole - 1u) / 2u;
        if (!(hl[parent] < vl))
            break;
        he[hole] = he[parent];
        hl[hole] = hl[parent];
        hole = parent;
    }
    he[hole] = ve;
    hl[hole] = vl;
}

// std::__adjust_heap: the hole sinks to a leaf along the larger child (the LEFT one only if the right one is smaller),
// then the value climbs back
LB_CHI_HD void chi_heap_adjust(uint32_t *he, double *hl, uint32_t hole, uint32_t len, uint32_t ve, double vl)
{
    const uint32_t top = hole;
    uint32_t child = hole;
    while (len >= 2u && child < (len - 1u) / 2u)
    {
        child = 2u * (child + 1u);
        if (hl[child] < hl[child - 1u])
            --child;
        he[hole] = he[child];
        hl[hole] = hl[child];
        hole = child;
    }
    if ((len & 1u) == 0u && len >= 2u && child == (len - 2u) / 2u)
    {
        child = 2u * (child + 1u);
        he[hole] = he[child - 1u];
        hl[hole] = hl[child - 1u];
        hole = child - 1u;
    }
    chi_heap_sift_up(he, hl, hole, top, ve, vl);
}

// emplace_back + std::push_heap
LB_CHI_HD void chi_heap_push(uint32_t *he, double *hl, uint32_t &size, uint32_t e, double len)
{
    chi_heap_sift_up(he, hl, size, 0u, e, len);
    ++size;
}

// std::pop_heap + back() + pop_back()
LB_CHI_HD void chi_heap_pop(uint32_t *he, double *hl, uint32_t &size, uint32_t &e, double &len)
{
    e = he[0];
    len = hl[0];
    --size;
    if (size >= 1u)
    {
        const uint32_t ve = he[size];
        const double vl = hl[size];
        chi_heap_adjust(he, hl, 0u, size, ve, vl);
    }
}

// ---- libstdc++ 13 std::sort(ids, by dist[i] < dist[j]) (delaunator.cpp:339-341), step by step ------------------------
// The elements travel as (key, id) records of 8 bytes, so that a comparison is one load and a cluster of 20 000 points
// sorts out of 160 KB; the key is the RANK of the point's distance among the distinct distances of the cluster (equal
// distances, equal ranks), which compares exactly like the distance itself. The sequence of comparisons and moves is the
// one std::sort performs on the ids.

struct alignas(8) ChiKeyed
{
    uint32_t d; // rank of dist[id] among the distinct distances
    uint32_t id;
};

LB_CHI_HD void chi_sort_adjust_heap(ChiKeyed *a, uint32_t hole, uint32_t len, ChiKeyed v)
{
    const uint32_t top = hole;
    uint32_t child = hole;
    while (len >= 2u && child < (len - 1u) / 2u)
    {
        child = 2u * (child + 1u);
        if (a[child].d < a[child - 1u].d)
            --child;
        a[hole] = a[child];
        hole = child;
    }
    if ((len & 1u) == 0u && len >= 2u && child == (len - 2u) / 2u)
    {
        child = 2u * (child + 1u);
        a[hole] = a[child - 1u];
        hole = child - 1u;
    }
    while (hole > top)
    {
        const uint32_t parent = (hole - 1u) / 2u;
        if (!(a[parent].d < v.d))
            break;
        a[hole] = a[parent];
        hole = parent;
    }
    a[hole] = v;
}

// std::__partial_sort(first, last, last): make_heap + sort_heap
LB_CHI_HD void chi_sort_heap_sort(ChiKeyed *a, uint32_t len)
{
    if (len < 2u)
        return;
    for (uint32_t parent = (len - 2u) / 2u;; --parent)
    {
        chi_sort_adjust_heap(a, parent, len, a[parent]);
        if (parent == 0u)
            break;
    }
    for (uint32_t last = len; last > 1u;)
    {
        --last;
        const ChiKeyed v = a[last];
        a[last] = a[0];
        chi_sort_adjust_heap(a, 0u, last, v);
    }
}

LB_CHI_HD void chi_sort_unguarded_insert(ChiKeyed *a, uint32_t last)
{
    const ChiKeyed v = a[last];
    uint32_t next = last - 1u;
    while (v.d < a[next].d)
    {
        a[last] = a[next];
        last = next;
        --next;
    }
    a[last] = v;
}

LB_CHI_HD void chi_sort_insertion_sort(ChiKeyed *a, uint32_t first, uint32_t last)
{
    if (first == last)
        return;
    for (uint32_t i = first + 1u; i != last; ++i)
    {
        if (a[i].d < a[first].d)
        {
            const ChiKeyed v = a[i];
            for (uint32_t j = i; j > first; --j)
                a[j] = a[j - 1u];
            a[first] = v;
        }
        else
            chi_sort_unguarded_insert(a, i);
    }
}

// __move_median_to_first(first, first + 1, mid, last - 1) + __unguarded_partition(first + 1, last, pivot = first):
// one step of std::__introsort_loop over [first, last), last - first > 16. Returns the cut.
LB_CHI_HD uint32_t chi_sort_partition_step(ChiKeyed *a, uint32_t first, uint32_t last)
{
    const uint32_t ia = first + 1u, ib = first + (last - first) / 2u, ic = last - 1u;
    const uint32_t ka = a[ia].d, kb = a[ib].d, kc = a[ic].d;
    uint32_t pick;
    if (ka < kb)
        pick = (kb < kc) ? ib : ((ka < kc) ? ic : ia);
    else
        pick = (ka < kc) ? ia : ((kb < kc) ? ic : ib);
    {
        const ChiKeyed t = a[first];
        a[first] = a[pick];
        a[pick] = t;
    }
    const uint32_t kp = a[first].d;
    uint32_t lo = first + 1u, hi = last;
    while (true)
    {
        while (a[lo].d < kp)
            ++lo;
        --hi;
        while (kp < a[hi].d)
            --hi;
        if (!(lo < hi))
            break;
        const ChiKeyed t = a[lo];
        a[lo] = a[hi];
        a[hi] = t;
        ++lo;
    }
    return lo;
}

// std::__introsort_loop(first, last, depth) with its recursion on the right part turned into a stack (the parts are
// disjoint, so the order in which they are finished does not show). Leaves runs of at most 16 elements unsorted.
LB_CHI_HD void chi_introsort_loop(ChiKeyed *a, uint32_t first0, uint32_t last0, uint32_t depth0)
{
    uint32_t st_first[64], st_last[64], st_depth[64];
    uint32_t sp = 1u;
    st_first[0] = first0;
    st_last[0] = last0;
    st_depth[0] = depth0;
    while (sp > 0u)
    {
        --sp;
        const uint32_t first = st_first[sp];
        uint32_t last = st_last[sp];
        uint32_t depth = st_depth[sp];
        while (last - first > 16u)
        {
            if (depth == 0u)
            {
                chi_sort_heap_sort(a + first, last - first);
                break;
            }
            --depth;
            const uint32_t cut = chi_sort_partition_step(a, first, last);
            // right part [cut, last) later, left part [first, cut) now
            if (sp < 64u)
            {
                st_first[sp] = cut;
                st_last[sp] = last;
                st_depth[sp] = depth;
                ++sp;
            }
            last = cut;
        }
    }
}

LB_CHI_HD uint32_t chi_sort_depth_limit(uint32_t n)
{
    uint32_t lg = 0u;
    for (uint32_t v = n; v > 1u; v >>= 1)
        ++lg;
    return 2u * lg;
}

// std::sort: introsort loop, then __final_insertion_sort. The latter is a stable insertion sort of what the loop left
// behind (its guarded and unguarded halves differ in bounds checks only), i.e. the result is that arrangement stably
// sorted by key - which is how the device finishes (chi_shape.cuh), with a parallel sort by (key, position).
LB_CHI_HD void chi_introsort(ChiKeyed *a, uint32_t n)
{
    if (n == 0u)
        return;
    chi_introsort_loop(a, 0u, n, chi_sort_depth_limit(n));
    if (n > 16u)
    {
        chi_sort_insertion_sort(a, 0u, 16u);
        for (uint32_t i = 16u; i != n; ++i)
            chi_sort_unguarded_insert(a, i);
    }
    else
        chi_sort_insertion_sort(a, 0u, n);
}

// ---- seed triangle (delaunator.cpp:214-327), sequential form ---------------------------------------------------------

// Bounding box, seed triangle, circumcentre, distances, hash keys. Returns kChiOk or the reason why the reference does
// not deliver. The device computes the same values with a warp (chi_shape.cuh); this is the definition.
LB_CHI_HD uint32_t chi_seed_sequential(ChiWork &w)
{
    const uint32_t n = w.n;
    double max_x = -DBL_MAX, max_y = -DBL_MAX, min_x = DBL_MAX, min_y = DBL_MAX;
    for (uint32_t i = 0; i < n; ++i)
    {
        const double x = chi_px(w, i), y = chi_py(w, i);
        min_x = (min_x < x) ? min_x : x; // std::min(p.x(), min_x)
        min_y = (min_y < y) ? min_y : y;
        max_x = (x < max_x) ? max_x : x; // std::max(p.x(), max_x)
        max_y = (y < max_y) ? max_y : y;
    }
    const double width = max_x - min_x;
    const double height = max_y - min_y;
    w.span = width * width + height * height;
    const double bx = (min_x + max_x) / 2.0, by = (min_y + max_y) / 2.0;
    uint32_t i0 = kChiNone, i1 = kChiNone, i2 = kChiNone;
    double best = DBL_MAX;
    for (uint32_t i = 0; i < n; ++i)
    {
        const double d = chi_dist2(chi_px(w, i), chi_py(w, i), bx, by);
        if (d < best)
        {
            i0 = i;
            best = d;
        }
    }
    if (i0 == kChiNone)
        return kChiErrCoincident;
    const double p0x = chi_px(w, i0), p0y = chi_py(w, i0);
    best = DBL_MAX;
    for (uint32_t i = 0; i < n; ++i)
    {
        if (i == i0)
            continue;
        const double d = chi_dist2(chi_px(w, i), chi_py(w, i), p0x, p0y);
        if (d < best && d > 0.0)
        {
            i1 = i;
            best = d;
        }
    }
    if (i1 == kChiNone)
        return kChiErrCoincident; // the reference dereferences m_points[INVALID_INDEX] here
    const double p1x = chi_px(w, i1), p1y = chi_py(w, i1);
    best = DBL_MAX;
    for (uint32_t i = 0; i < n; ++i)
    {
        if (i == i0 || i == i1)
            continue;
        const double r = chi_circumradius2(p0x, p0y, p1x, p1y, chi_px(w, i), chi_py(w, i));
        if (r < best)
        {
            i2 = i;
            best = r;
        }
    }
    if (!(best < DBL_MAX))
        return kChiErrCollinear; // throw std::runtime_error("not triangulation")
    if (chi_ccw(p0x, p0y, p1x, p1y, chi_px(w, i2), chi_py(w, i2)))
    {
        const uint32_t t = i1;
        i1 = i2;
        i2 = t;
    }
    w.i0 = i0;
    w.i1 = i1;
    w.i2 = i2;
    w.s0x = p0x;
    w.s0y = p0y;
    w.s1x = chi_px(w, i1);
    w.s1y = chi_py(w, i1);
    w.s2x = chi_px(w, i2);
    w.s2y = chi_py(w, i2);
    chi_circumcentre(w.s0x, w.s0y, w.s1x, w.s1y, w.s2x, w.s2y, w.cx, w.cy);
    for (uint32_t i = 0; i < n; ++i)
    {
        w.dist[i] = chi_dist2(chi_px(w, i), chi_py(w, i), w.cx, w.cy);
        w.node[i].key = chi_hash_key(w, chi_px(w, i), chi_py(w, i));
    }
    return kChiOk;
}

// the point is skipped as "one of the seed triangle's" (delaunator.cpp:394-398)
LB_CHI_HD bool chi_on_seed(const ChiWork &w, double x, double y)
{
    return chi_same(x, y, w.s0x, w.s0y) || chi_same(x, y, w.s1x, w.s1y) || chi_same(x, y, w.s2x, w.s2y);
}

// ---- triangulation (delaunator.cpp:343-487, 520-685) ------------------------------------------------------------------

LB_CHI_HD uint32_t chi_next_half_edge(uint32_t e)
{
    return (e % 3u == 2u) ? e - 2u : e + 1u;
}

// a triangle corner as the sweep hands it around: vertex id and coordinates
struct ChiCorner
{
    uint32_t v;
    float x, y;
};

LB_CHI_HD ChiCorner chi_corner(uint32_t v, double x, double y)
{
    ChiCorner c;
    c.v = v;
    c.x = static_cast<float>(x); // exact: the coordinates are widened floats
    c.y = static_cast<float>(y);
    return c;
}

// Delaunator::addTriangle: three records, three links. Returns the first half-edge; ea/eb/ec receive the records as
// written (the legalisation that follows starts from them without reading them back).
LB_CHI_HD uint32_t chi_add_triangle(ChiWork &w, const ChiCorner &p0, const ChiCorner &p1, const ChiCorner &p2, uint32_t a,
                                    uint32_t b, uint32_t c, ChiEdge &ea, ChiEdge &eb, ChiEdge &ec)
{
    uint32_t t = w.n_half;
    if (t + 3u > 3u * (2u * w.n - 5u)) // (a triangulation of n points has at most 2n - 5 triangles; only corrupt links get here)
    {
        w.overflow = true;
        t -= 3u;
    }
    w.n_half = t + 3u;
    ea.twin = a;
    ea.v = p0.v;
    ea.x = p0.x;
    ea.y = p0.y;
    eb.twin = b;
    eb.v = p1.v;
    eb.x = p1.x;
    eb.y = p1.y;
    ec.twin = c;
    ec.v = p2.v;
    ec.x = p2.x;
    ec.y = p2.y;
    chi_store_edge(&w.edge[t], ea);
    chi_store_edge(&w.edge[t + 1u], eb);
    chi_store_edge(&w.edge[t + 2u], ec);
    if (a != kChiNone)
        w.edge[a].twin = t;
    if (b != kChiNone)
        w.edge[b].twin = t + 1u;
    if (c != kChiNone)
        w.edge[c].twin = t + 2u;
    return t;
}

// Delaunator::legalize: flips until the pair of triangles across every touched edge is locally Delaunay. `have` says
// that e_a / e_al / e_ar already hold the records of a and of the two other half-edges of its triangle. After a flip
// the same half-edge is looked at again (delaunator.cpp:520-633) and its triangle's records are known without a load;
// a half-edge popped from the stack reloads its triangle.
LB_CHI_HD uint32_t chi_legalize(ChiWork &w, uint32_t a, ChiEdge e_a, ChiEdge e_al, ChiEdge e_ar, bool have, uint32_t *spill,
                                uint32_t spill_cap, uint32_t &guard)
{
    uint32_t local_stack[kChiLocalStack];
    uint32_t depth = 0u;
    uint32_t ar = 0u;
    while (true)
    {
        const uint32_t ra = a % 3u;
        const uint32_t a0 = a - ra;
        ar = a0 + (ra == 0u ? 2u : ra - 1u);              // a0 + (a + 2) % 3
        const uint32_t al = a0 + (ra == 2u ? 0u : ra + 1u); // a0 + (a + 1) % 3
        if (!have)
        {
            e_a = chi_load_edge(&w.edge[a]);
            e_al = chi_load_edge(&w.edge[al]);
            e_ar = chi_load_edge(&w.edge[ar]);
            have = true;
        }
        const uint32_t b = e_a.twin;
        bool flipped = false;
        if (b != kChiNone)
        {
            const uint32_t rb = b % 3u;
            const uint32_t b0 = b - rb;
            const uint32_t bl = b0 + (rb == 0u ? 2u : rb - 1u); // b0 + (b + 2) % 3
            const ChiEdge e_bl = chi_load_edge(&w.edge[bl]);
            // p0 = triangles[ar], pr = triangles[a], pl = triangles[al], p1 = triangles[bl]
            if (chi_in_circle(static_cast<double>(e_ar.x), static_cast<double>(e_ar.y), static_cast<double>(e_a.x),
                              static_cast<double>(e_a.y), static_cast<double>(e_al.x), static_cast<double>(e_al.y),
                              static_cast<double>(e_bl.x), static_cast<double>(e_bl.y)))
            {
                const uint32_t hbl = e_bl.twin;
                if (hbl == kChiNone)
                {
                    // the flipped edge was a hull edge: its hull_tri entry follows it (delaunator.cpp:592-606)
                    uint32_t e = w.hull_start, steps = 0u;
                    do
                    {
                        if (w.node[e].tri == bl)
                        {
                            w.node[e].tri = a;
                            w.tri_moved = true;
                            break;
                        }
                        e = w.node[e].prev;
                    } while (e != w.hull_start && ++steps <= w.n);
                }
                const uint32_t har = e_ar.twin;
                // triangles[a] = p1, link(a, hbl)
                e_a.v = e_bl.v;
                e_a.x = e_bl.x;
                e_a.y = e_bl.y;
                e_a.twin = hbl;
                chi_store_edge(&w.edge[a], e_a);
                if (hbl != kChiNone)
                    w.edge[hbl].twin = a;
                // triangles[b] = p0, link(b, half_edges[ar])
                ChiEdge e_b;
                e_b.v = e_ar.v;
                e_b.x = e_ar.x;
                e_b.y = e_ar.y;
                e_b.twin = har;
                chi_store_edge(&w.edge[b], e_b);
                if (har != kChiNone)
                    w.edge[har].twin = b;
                // link(ar, bl)
                e_ar.twin = bl;
                w.edge[ar].twin = bl;
                w.edge[bl].twin = ar;
                if (guard == 0u || (depth >= kChiLocalStack && depth - kChiLocalStack >= spill_cap))
                {
                    guard = 0u;
                    return ar;
                }
                --guard;
                const uint32_t br = b0 + (rb == 2u ? 0u : rb + 1u); // b0 + (b + 1) % 3
                if (depth < kChiLocalStack)
                    local_stack[depth] = br;
                else
                    spill[depth - kChiLocalStack] = br;
                ++depth;
                flipped = true; // e_a, e_al, e_ar describe the triangle of `a` after the flip
            }
        }
        if (!flipped)
        {
            if (depth == 0u)
                break;
            --depth;
            a = depth < kChiLocalStack ? local_stack[depth] : spill[depth - kChiLocalStack];
            have = false;
        }
    }
    return ar;
}

// The advancing-hull sweep over ids[0..n) (already in the reference's order). Returns kChiOk / kChiErrGuard.
LB_CHI_HD uint32_t chi_triangulate(ChiWork &w)
{
    const uint32_t n = w.n;
    uint32_t *spill = reinterpret_cast<uint32_t *>(w.dist); // the distances are spent once the order stands
    const uint32_t spill_cap = 2u * n;
    uint32_t guard = 0xFFFFFFF0u; // (flips are finite for the reference as well; a budget keeps a corrupt input from hanging the GPU)
    if (static_cast<unsigned long long>(n) * 64ull < guard)
        guard = n * 64u;
    for (uint32_t h = 0; h < w.hash_size; ++h)
        w.hash[h] = kChiNone;
    const uint32_t i0 = w.i0, i1 = w.i1, i2 = w.i2;
    w.hull_start = i0;
    w.node[i0].next = w.node[i2].prev = i1;
    w.node[i1].next = w.node[i0].prev = i2;
    w.node[i2].next = w.node[i1].prev = i0;
    w.node[i0].tri = 0u;
    w.node[i1].tri = 1u;
    w.node[i2].tri = 2u;
    w.hash[w.node[i0].key] = i0;
    w.hash[w.node[i1].key] = i1;
    w.hash[w.node[i2].key] = i2;
    w.n_half = 0u;
    ChiEdge ea, eb, ec;
    chi_add_triangle(w, chi_corner(i0, w.s0x, w.s0y), chi_corner(i1, w.s1x, w.s1y), chi_corner(i2, w.s2x, w.s2y), kChiNone,
                     kChiNone, kChiNone, ea, eb, ec);
    double xp = 0.0, yp = 0.0;
    const double span_4eps = 4.0 * kChiEps * w.span;
    // the next point's record is fetched while the current one is being inserted (its coordinates and key never change;
    // prev / next / tri of a point that is not on the hull yet are not read)
    ChiNode nd_next = chi_load_node(&w.node[w.ids[0]]);
    uint32_t i_next = w.ids[0];
    for (uint32_t k = 0; k < n; ++k)
    {
        const uint32_t i = i_next;
        const ChiNode nd_i = nd_next;
        if (k + 1u < n)
        {
            i_next = w.ids[k + 1u];
            nd_next = chi_load_node(&w.node[i_next]);
        }
        const double x = nd_i.x, y = nd_i.y;
        if (k > 0u && chi_same(x, y, xp, yp))
            continue;
        xp = x;
        yp = y;
        if (chi_on_seed(w, x, y))
            continue;
        // a hull vertex near the point's direction, from the pseudo-angle hash
        uint32_t start = 0u;
        const uint32_t key = nd_i.key;
        ChiNode nd_s;
        bool found = false;
        for (uint32_t j = 0; j < w.hash_size && !found; ++j)
        {
            uint32_t slot = key + j;
            if (slot >= w.hash_size)
                slot %= w.hash_size;
            start = w.hash[slot];
            if (start != kChiNone)
            {
                nd_s = chi_load_node(&w.node[start]);
                found = start != nd_s.next; // (a vertex that left the hull points at itself)
            }
        }
        if (!found)
            return kChiErrGuard; // (not reachable: the hull's last two insertions are always in the hash)
        start = nd_s.prev;
        uint32_t e = start, q, steps = 0u;
        ChiNode nd_e = chi_load_node(&w.node[e]);
        ChiNode nd_q;
        while (true)
        {
            if (++steps > n + 1u)
                return kChiErrGuard;
            q = nd_e.next;
            nd_q = chi_load_node(&w.node[q]);
            // Point::equal(p, hull vertex, span): squared distance / span < epsilon
            if (chi_near(chi_dist2(nd_e.x, nd_e.y, x, y), w.span, span_4eps) ||
                chi_near(chi_dist2(nd_q.x, nd_q.y, x, y), w.span, span_4eps))
            {
                e = kChiNone;
                break;
            }
            if (chi_ccw(x, y, nd_e.x, nd_e.y, nd_q.x, nd_q.y))
                break;
            e = q;
            nd_e = nd_q;
            if (e == start)
            {
                e = kChiNone;
                break;
            }
        }
        if (e == kChiNone)
            continue;
        const ChiCorner c_i = chi_corner(i, x, y);
        // add the first triangle from the point: (e, i, hull_next[e]) against hull_tri[e]
        w.tri_moved = false;
        uint32_t t = chi_add_triangle(w, chi_corner(e, nd_e.x, nd_e.y), c_i, chi_corner(q, nd_q.x, nd_q.y), kChiNone, kChiNone,
                                      nd_e.tri, ea, eb, ec);
        uint32_t tri_i = chi_legalize(w, t + 2u, ec, ea, eb, true, spill, spill_cap, guard); // hull_tri[i]
        uint32_t tri_e = t;                                                                   // hull_tri[e] = t
        w.node[e].tri = t;
        // walk forward through the hull, adding more triangles and flipping
        uint32_t next = q;
        ChiNode nd_n = nd_q;
        while (true)
        {
            q = nd_n.next;
            nd_q = chi_load_node(&w.node[q]);
            if (w.overflow || !chi_ccw(x, y, nd_n.x, nd_n.y, nd_q.x, nd_q.y))
                break;
            const uint32_t tri_n = w.tri_moved ? w.node[next].tri : nd_n.tri;
            t = chi_add_triangle(w, chi_corner(next, nd_n.x, nd_n.y), c_i, chi_corner(q, nd_q.x, nd_q.y), tri_i, kChiNone, tri_n,
                                 ea, eb, ec);
            tri_i = chi_legalize(w, t + 2u, ec, ea, eb, true, spill, spill_cap, guard);
            w.node[next].next = next; // mark as removed
            next = q;
            nd_n = nd_q;
        }
        // walk backward from the other side, adding more triangles and flipping
        if (e == start)
        {
            while (true)
            {
                q = nd_e.prev;
                nd_q = chi_load_node(&w.node[q]);
                if (w.overflow || !chi_ccw(x, y, nd_q.x, nd_q.y, nd_e.x, nd_e.y))
                    break;
                const uint32_t tri_q = w.tri_moved ? w.node[q].tri : nd_q.tri;
                if (w.tri_moved)
                    tri_e = w.node[e].tri;
                t = chi_add_triangle(w, chi_corner(q, nd_q.x, nd_q.y), c_i, chi_corner(e, nd_e.x, nd_e.y), kChiNone, tri_e, tri_q,
                                     ea, eb, ec);
                chi_legalize(w, t + 2u, ec, ea, eb, true, spill, spill_cap, guard);
                w.node[q].tri = t;
                tri_e = t; // (e becomes q below)
                w.node[e].next = e; // mark as removed
                e = q;
                nd_e = nd_q;
            }
        }
        // update the hull indices
        // (hull_tri[i] lived in a register so far: i is not on the hull before this point, no flip can have moved it)
        w.node[i].prev = e;
        w.node[i].next = next;
        w.node[i].tri = tri_i;
        w.hull_start = e;
        w.node[next].prev = i;
        w.node[e].next = i;
        w.hash[key] = i;
        w.hash[nd_e.key] = e;
        if (guard == 0u || w.overflow)
            return kChiErrGuard;
    }
    return kChiOk;
}

// ---- erosion of the boundary (concave_hull.hpp:96-193) -----------------------------------------------------------------

// Delaunator::edgeLength (delaunator.cpp:709-720); std::pow(v, 2.0) is v * v in the reference binary
LB_CHI_HD double chi_edge_length(const ChiEdge &a, const ChiEdge &b)
{
    const double dx = static_cast<double>(a.x) - static_cast<double>(b.x);
    const double dy = static_cast<double>(a.y) - static_cast<double>(b.y);
    return sqrt(dx * dx + dy * dy);
}

LB_CHI_HD double chi_edge_length(const ChiWork &w, uint32_t e)
{
    return chi_edge_length(chi_load_edge(&w.edge[e]), chi_load_edge(&w.edge[chi_next_half_edge(e)]));
}

// Erodes the hull in place (hull_next / hull_prev) and returns the number of vertices of the closed outline, i.e.
// the length of getHullIndices() with the start repeated; out[] receives the open loop (at most n entries).
LB_CHI_HD uint32_t chi_erode_and_walk(ChiWork &w, uint32_t *out, bool onb_cleared = false)
{
    const uint32_t n = w.n;
    uint32_t *he = w.ids;
    double *hl = w.dist;
    uint32_t size = 0u;
    if (!onb_cleared)
        for (uint32_t i = 0; i < n; ++i)
            w.onb[i] = 0u;
    double max_len = -DBL_MAX, min_len = DBL_MAX;
    // boundary_indices = getHullIndices(): the loop from hull_start plus hull_start once more, so the first edge is
    // in the heap twice
    uint32_t v = w.hull_start;
    bool closing = false;
    while (true)
    {
        w.onb[v] = 1u;
        const uint32_t e = w.node[v].tri;
        const double len = chi_edge_length(w, e);
        chi_heap_push(he, hl, size, e, len);
        min_len = (min_len < len) ? min_len : len; // std::min(len, min_len)
        max_len = (len < max_len) ? max_len : len; // std::max(len, max_len)
        if (closing)
            break;
        v = w.node[v].next;
        if (v == w.hull_start)
            closing = true;
    }
    const double length_param = kChiFactor * max_len + (1.0 - kChiFactor) * min_len;
    while (size > 0u)
    {
        uint32_t e;
        double len;
        chi_heap_pop(he, hl, size, e, len);
        if (len <= length_param)
            break;
        const uint32_t e_n = chi_next_half_edge(e);
        const uint32_t e_p = chi_next_half_edge(e_n);
        const ChiEdge r_e = chi_load_edge(&w.edge[e]), r_n = chi_load_edge(&w.edge[e_n]), r_p = chi_load_edge(&w.edge[e_p]);
        const uint32_t c = r_p.v; // getInteriorPoint
        if (w.onb[c])
            continue;
        const uint32_t e_b = r_n.twin;
        const uint32_t e_a = r_p.twin;
        if (e_a == kChiNone || e_b == kChiNone || size + 2u > n + 1u)
            break; // (not reachable: an interior vertex has no hull edge, and every erosion adds one vertex)
        const double len_a = chi_edge_length(w, e_a);
        const double len_b = chi_edge_length(w, e_b);
        chi_heap_push(he, hl, size, e_a, len_a);
        chi_heap_push(he, hl, size, e_b, len_b);
        const uint32_t a = r_e.v;
        const uint32_t b = r_n.v;
        w.node[c].next = b;
        w.node[c].prev = a;
        w.node[a].next = c;
        w.node[b].prev = c;
        w.onb[c] = 1u;
    }
    uint32_t h = 0u;
    v = w.hull_start;
    do
    {
        if (h < n)
            out[h] = v;
        ++h;
        v = w.node[v].next;
    } while (v != w.hull_start && h <= n);
    return h + 1u;
}

} // namespace lb
