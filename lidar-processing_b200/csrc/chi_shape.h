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

struct alignas(16) ChiXY // the coordinates as the reference's delaunator sees them: static_cast<double>(float)
{
    double x, y;
};

// Per-cluster working set. Every array lives in one block of chi_layout(n).bytes bytes.
struct ChiWork
{
    const ChiXY *xy;  // [n] the cluster's points (x, y of the grouped PointXYZ records, widened once)
    uint16_t *key;    // [n] getHashKey of every point (it depends on the point and the seed circumcentre only)
    uint32_t n;
    double *dist;     // [n + 1] squared distance from the seed circumcentre; later: legalisation stack, then heap lengths
    uint32_t *ids;    // [n + 1] point order; later: heap edges
    uint32_t *tri;    // [3 * (2n - 5)] triangles
    uint32_t *half;   // [3 * (2n - 5)] half_edges
    uint32_t *hprev, *hnext, *htri; // [n] advancing hull
    uint32_t *hash;   // [hash_size]
    uint8_t *onb;     // [n] boundary_set of concave_hull.hpp:110
    uint32_t hash_size;
    uint32_t n_half;  // triangles.size()
    bool overflow;    // guard: more triangles than a triangulation can have
    uint32_t hull_start;
    uint32_t i0, i1, i2;
    double cx, cy;    // m_center
    double span;
    double s0x, s0y, s1x, s1y, s2x, s2y; // seed triangle coordinates (after the orientation swap)
};

struct ChiLayout
{
    size_t dist, ids, tri, half, hprev, hnext, htri, hash, onb, xy, key, bytes;
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

LB_CHI_HD ChiLayout chi_layout(uint32_t n)
{
    ChiLayout l;
    const size_t np = n, t3 = n >= 3u ? 3u * (2u * np - 5u) : 3u;
    l.hash_size = chi_hash_size(n);
    size_t at = 0;
    l.dist = at;
    at = chi_align16(at + 8u * (np + 1u));
    l.ids = at;
    at = chi_align16(at + 4u * (np + 1u));
    l.tri = at;
    at = chi_align16(at + 4u * t3);
    l.half = at;
    at = chi_align16(at + 4u * t3);
    l.hprev = at;
    at = chi_align16(at + 4u * np);
    l.hnext = at;
    at = chi_align16(at + 4u * np);
    l.htri = at;
    at = chi_align16(at + 4u * np);
    l.hash = at;
    at = chi_align16(at + 4u * l.hash_size);
    l.onb = at;
    at = chi_align16(at + np);
    l.xy = at;
    at = chi_align16(at + 16u * np);
    l.key = at;
    at = chi_align16(at + 2u * np);
    l.bytes = at;
    return l;
}

LB_CHI_HD void chi_bind(ChiWork &w, unsigned char *block, const ChiLayout &l, uint32_t n)
{
    w.n = n;
    w.dist = reinterpret_cast<double *>(block + l.dist);
    w.ids = reinterpret_cast<uint32_t *>(block + l.ids);
    w.tri = reinterpret_cast<uint32_t *>(block + l.tri);
    w.half = reinterpret_cast<uint32_t *>(block + l.half);
    w.hprev = reinterpret_cast<uint32_t *>(block + l.hprev);
    w.hnext = reinterpret_cast<uint32_t *>(block + l.hnext);
    w.htri = reinterpret_cast<uint32_t *>(block + l.htri);
    w.hash = reinterpret_cast<uint32_t *>(block + l.hash);
    w.onb = block + l.onb;
    w.xy = reinterpret_cast<const ChiXY *>(block + l.xy);
    w.key = reinterpret_cast<uint16_t *>(block + l.key);
    w.hash_size = l.hash_size;
    w.n_half = 0u;
    w.overflow = false;
}

LB_CHI_HD double chi_px(const ChiWork &w, uint32_t i)
{
    return w.xy[i].x;
}

LB_CHI_HD double chi_py(const ChiWork &w, uint32_t i)
{
    return w.xy[i].y;
}

// both coordinates with one 16-byte load
LB_CHI_HD void chi_pt(const ChiWork &w, uint32_t i, double &x, double &y)
{
#ifdef __CUDA_ARCH__
    const double2 p = *reinterpret_cast<const double2 *>(&w.xy[i]);
#else
    const ChiXY p = w.xy[i];
#endif
    x = p.x;
    y = p.y;
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
// The elements travel as (distance, id) records, so that a comparison is one load; the sequence of comparisons and moves
// is the one std::sort performs on the ids.

struct alignas(16) ChiKeyed
{
    double d;
    uint32_t id;
    uint32_t pad;
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

LB_CHI_HD void chi_introsort(ChiKeyed *a, uint32_t n)
{
    if (n == 0u)
        return;
    // std::__introsort_loop with its recursion on the right part turned into a stack (the parts are disjoint, so the
    // order in which they are finished does not show)
    uint32_t st_first[64], st_last[64], st_depth[64];
    uint32_t sp = 0u;
    uint32_t lg = 0u;
    for (uint32_t v = n; v > 1u; v >>= 1)
        ++lg;
    st_first[0] = 0u;
    st_last[0] = n;
    st_depth[0] = 2u * lg;
    sp = 1u;
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
            // __move_median_to_first(first, first + 1, mid, last - 1)
            const uint32_t ia = first + 1u, ib = first + (last - first) / 2u, ic = last - 1u;
            const double ka = a[ia].d, kb = a[ib].d, kc = a[ic].d;
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
            // __unguarded_partition(first + 1, last, pivot = first)
            const double kp = a[first].d;
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
            // right part [lo, last) later, left part [first, lo) now
            if (sp < 64u)
            {
                st_first[sp] = lo;
                st_last[sp] = last;
                st_depth[sp] = depth;
                ++sp;
            }
            last = lo;
        }
    }
    // __final_insertion_sort
    if (n > 16u)
    {
        chi_sort_insertion_sort(a, 0u, 16u);
        for (uint32_t i = 16u; i != n; ++i)
            chi_sort_unguarded_insert(a, i);
    }
    else
        chi_sort_insertion_sort(a, 0u, n);
}

// ids[0..n) <- what std::sort leaves of 0..n-1 under dist[i] < dist[j]; `scratch` holds n ChiKeyed records
LB_CHI_HD void chi_introsort_ids(uint32_t *ids, const double *dist, uint32_t n, ChiKeyed *scratch)
{
    for (uint32_t i = 0; i < n; ++i)
    {
        scratch[i].d = dist[i];
        scratch[i].id = i;
        scratch[i].pad = 0u;
    }
    chi_introsort(scratch, n);
    for (uint32_t i = 0; i < n; ++i)
        ids[i] = scratch[i].id;
}

// ---- seed triangle (delaunator.cpp:214-327), sequential form ---------------------------------------------------------

// Bounding box, seed triangle, circumcentre, distances. Returns kChiOk or the reason why the reference does not
// deliver. The device computes the same values with a warp (chi_shape.cuh); this is the definition.
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
        w.key[i] = static_cast<uint16_t>(chi_hash_key(w, chi_px(w, i), chi_py(w, i)));
    }
    return kChiOk;
}

// the point is skipped as "one of the seed triangle's" (delaunator.cpp:394-398)
LB_CHI_HD bool chi_on_seed(const ChiWork &w, double x, double y)
{
    return chi_same(x, y, w.s0x, w.s0y) || chi_same(x, y, w.s1x, w.s1y) || chi_same(x, y, w.s2x, w.s2y);
}

// ---- triangulation (delaunator.cpp:343-487, 520-685) ------------------------------------------------------------------

LB_CHI_HD void chi_link(ChiWork &w, uint32_t a, uint32_t b)
{
    w.half[a] = b;
    if (b != kChiNone)
        w.half[b] = a;
}

LB_CHI_HD uint32_t chi_add_triangle(ChiWork &w, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t t = w.n_half;
    if (t + 3u > 3u * (2u * w.n - 5u)) // (a triangulation of n points has at most 2n - 5 triangles; only corrupt links get here)
    {
        w.overflow = true;
        t -= 3u;
        w.n_half = t;
    }
    w.tri[t] = p0;
    w.tri[t + 1u] = p1;
    w.tri[t + 2u] = p2;
    w.n_half = t + 3u;
    chi_link(w, t, a);
    chi_link(w, t + 1u, b);
    chi_link(w, t + 2u, c);
    return t;
}

// Delaunator::legalize: flips until the pair of triangles across every touched edge is locally Delaunay. The pending
// edges wait in `stack` (capacity stack_cap; the reference's vector grows without bound).
LB_CHI_HD uint32_t chi_legalize(ChiWork &w, uint32_t a, uint32_t *stack, uint32_t stack_cap, uint32_t &guard)
{
    uint32_t depth = 0u;
    uint32_t ar = 0u;
    while (true)
    {
        const uint32_t b = w.half[a];
        const uint32_t ra = a % 3u;
        const uint32_t a0 = a - ra;
        ar = a0 + (ra == 0u ? 2u : ra - 1u); // a0 + (a + 2) % 3
        bool flipped = false;
        if (b != kChiNone)
        {
            const uint32_t rb = b % 3u;
            const uint32_t b0 = b - rb;
            const uint32_t al = a0 + (ra == 2u ? 0u : ra + 1u); // a0 + (a + 1) % 3
            const uint32_t bl = b0 + (rb == 0u ? 2u : rb - 1u); // b0 + (b + 2) % 3
            const uint32_t p0 = w.tri[ar], pr = w.tri[a], pl = w.tri[al], p1 = w.tri[bl];
            double p0x, p0y, prx, pry, plx, ply, p1x, p1y;
            chi_pt(w, p0, p0x, p0y);
            chi_pt(w, pr, prx, pry);
            chi_pt(w, pl, plx, ply);
            chi_pt(w, p1, p1x, p1y);
            if (chi_in_circle(p0x, p0y, prx, pry, plx, ply, p1x, p1y))
            {
                w.tri[a] = p1;
                w.tri[b] = p0;
                const uint32_t hbl = w.half[bl];
                if (hbl == kChiNone)
                {
                    // the flipped edge was a hull edge: its hull_tri entry follows it (delaunator.cpp:592-606)
                    uint32_t e = w.hull_start, steps = 0u;
                    do
                    {
                        if (w.htri[e] == bl)
                        {
                            w.htri[e] = a;
                            break;
                        }
                        e = w.hprev[e];
                    } while (e != w.hull_start && ++steps <= w.n);
                }
                chi_link(w, a, hbl);
                chi_link(w, b, w.half[ar]);
                chi_link(w, ar, bl);
                if (depth >= stack_cap || guard == 0u)
                {
                    guard = 0u;
                    return ar;
                }
                --guard;
                stack[depth++] = b0 + (rb == 2u ? 0u : rb + 1u); // br = b0 + (b + 1) % 3
                flipped = true;
            }
        }
        if (!flipped)
        {
            if (depth == 0u)
                break;
            a = stack[--depth];
        }
    }
    return ar;
}

// The advancing-hull sweep over ids[0..n) (already in the reference's order). Returns kChiOk / kChiErrGuard.
LB_CHI_HD uint32_t chi_triangulate(ChiWork &w)
{
    const uint32_t n = w.n;
    uint32_t *stack = reinterpret_cast<uint32_t *>(w.dist); // the distances are spent once the order stands
    const uint32_t stack_cap = 2u * n;
    uint32_t guard = 0xFFFFFFF0u; // (flips are finite for the reference as well; a budget keeps a corrupt input from hanging the GPU)
    if (static_cast<unsigned long long>(n) * 64ull < guard)
        guard = n * 64u;
    for (uint32_t h = 0; h < w.hash_size; ++h)
        w.hash[h] = kChiNone;
    const uint32_t i0 = w.i0, i1 = w.i1, i2 = w.i2;
    w.hull_start = i0;
    w.hnext[i0] = w.hprev[i2] = i1;
    w.hnext[i1] = w.hprev[i0] = i2;
    w.hnext[i2] = w.hprev[i1] = i0;
    w.htri[i0] = 0u;
    w.htri[i1] = 1u;
    w.htri[i2] = 2u;
    w.hash[w.key[i0]] = i0;
    w.hash[w.key[i1]] = i1;
    w.hash[w.key[i2]] = i2;
    w.n_half = 0u;
    chi_add_triangle(w, i0, i1, i2, kChiNone, kChiNone, kChiNone);
    double xp = 0.0, yp = 0.0;
    const double span_4eps = 4.0 * kChiEps * w.span;
    for (uint32_t k = 0; k < n; ++k)
    {
        const uint32_t i = w.ids[k];
        double x, y;
        chi_pt(w, i, x, y);
        if (k > 0u && chi_same(x, y, xp, yp))
            continue;
        xp = x;
        yp = y;
        if (chi_on_seed(w, x, y))
            continue;
        // a hull vertex near the point's direction, from the pseudo-angle hash
        uint32_t start = 0u;
        const uint32_t key = w.key[i];
        for (uint32_t j = 0; j < w.hash_size; ++j)
        {
            uint32_t slot = key + j;
            if (slot >= w.hash_size)
                slot %= w.hash_size;
            start = w.hash[slot];
            if (start != kChiNone && start != w.hnext[start])
                break;
        }
        if (start == kChiNone)
            return kChiErrGuard; // (not reachable: the hull's last two insertions are always in the hash)
        start = w.hprev[start];
        uint32_t e = start, q, steps = 0u;
        while (true)
        {
            if (++steps > n + 1u)
                return kChiErrGuard;
            q = w.hnext[e];
            double ex, ey, qx, qy;
            chi_pt(w, e, ex, ey);
            chi_pt(w, q, qx, qy);
            // Point::equal(p, hull vertex, span): squared distance / span < epsilon
            if (chi_near(chi_dist2(ex, ey, x, y), w.span, span_4eps) || chi_near(chi_dist2(qx, qy, x, y), w.span, span_4eps))
            {
                e = kChiNone;
                break;
            }
            if (chi_ccw(x, y, ex, ey, qx, qy))
                break;
            e = q;
            if (e == start)
            {
                e = kChiNone;
                break;
            }
        }
        if (e == kChiNone)
            continue;
        uint32_t t = chi_add_triangle(w, e, i, w.hnext[e], kChiNone, kChiNone, w.htri[e]);
        w.htri[i] = chi_legalize(w, t + 2u, stack, stack_cap, guard);
        w.htri[e] = t;
        uint32_t next = w.hnext[e];
        while (true)
        {
            q = w.hnext[next];
            double nx, ny, qx, qy;
            chi_pt(w, next, nx, ny);
            chi_pt(w, q, qx, qy);
            if (w.overflow || !chi_ccw(x, y, nx, ny, qx, qy))
                break;
            t = chi_add_triangle(w, next, i, q, w.htri[i], kChiNone, w.htri[next]);
            w.htri[i] = chi_legalize(w, t + 2u, stack, stack_cap, guard);
            w.hnext[next] = next;
            next = q;
        }
        if (e == start)
        {
            while (true)
            {
                q = w.hprev[e];
                double ex, ey, qx, qy;
                chi_pt(w, e, ex, ey);
                chi_pt(w, q, qx, qy);
                if (w.overflow || !chi_ccw(x, y, qx, qy, ex, ey))
                    break;
                t = chi_add_triangle(w, q, i, e, kChiNone, w.htri[e], w.htri[q]);
                chi_legalize(w, t + 2u, stack, stack_cap, guard);
                w.htri[q] = t;
                w.hnext[e] = e;
                e = q;
            }
        }
        w.hprev[i] = e;
        w.hull_start = e;
        w.hprev[next] = i;
        w.hnext[e] = i;
        w.hnext[i] = next;
        w.hash[key] = i;
        w.hash[w.key[e]] = e;
        if (guard == 0u || w.overflow)
            return kChiErrGuard;
    }
    return kChiOk;
}

// ---- erosion of the boundary (concave_hull.hpp:96-193) -----------------------------------------------------------------

LB_CHI_HD uint32_t chi_next_half_edge(uint32_t e)
{
    return (e % 3u == 2u) ? e - 2u : e + 1u;
}

// Delaunator::edgeLength (delaunator.cpp:709-720); std::pow(v, 2.0) is v * v in the reference binary
LB_CHI_HD double chi_edge_length(const ChiWork &w, uint32_t e)
{
    const uint32_t a = w.tri[e], b = w.tri[chi_next_half_edge(e)];
    const double dx = chi_px(w, a) - chi_px(w, b);
    const double dy = chi_py(w, a) - chi_py(w, b);
    return sqrt(dx * dx + dy * dy);
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
        const uint32_t e = w.htri[v];
        const double len = chi_edge_length(w, e);
        chi_heap_push(he, hl, size, e, len);
        min_len = (min_len < len) ? min_len : len; // std::min(len, min_len)
        max_len = (len < max_len) ? max_len : len; // std::max(len, max_len)
        if (closing)
            break;
        v = w.hnext[v];
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
        const uint32_t c = w.tri[chi_next_half_edge(e_n)]; // getInteriorPoint
        if (w.onb[c])
            continue;
        const uint32_t e_b = w.half[e_n];
        const uint32_t e_a = w.half[chi_next_half_edge(e_n)];
        if (e_a == kChiNone || e_b == kChiNone || size + 2u > n + 1u)
            break; // (not reachable: an interior vertex has no hull edge, and every erosion adds one vertex)
        const double len_a = chi_edge_length(w, e_a);
        const double len_b = chi_edge_length(w, e_b);
        chi_heap_push(he, hl, size, e_a, len_a);
        chi_heap_push(he, hl, size, e_b, len_b);
        const uint32_t a = w.tri[e];
        const uint32_t b = w.tri[e_n];
        w.hnext[c] = b;
        w.hprev[c] = a;
        w.hnext[a] = c;
        w.hprev[b] = c;
        w.onb[c] = 1u;
    }
    uint32_t h = 0u;
    v = w.hull_start;
    do
    {
        if (h < n)
            out[h] = v;
        ++h;
        v = w.hnext[v];
    } while (v != w.hull_start && h <= n);
    return h + 1u;
}

} // namespace lb
