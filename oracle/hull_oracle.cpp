// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's ordered convex outlines
// (SURVEY.md §8f row 3). Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
// Pinned against the UNMODIFIED reference (oracle/_ref/libref_hull.so, built from
// /root/reference/src/polygon_simplification.cpp) by tests/test_oracle_pinning.py on the clusters of
// the golden frames and on synthetic clusters with duplicates and collinear runs.
//
// Follows: geom::Point<float>::operator< / operator==        reference Convex-Hull/convex_hull.hpp:51-73
//          crossProduct / getOrientation                       convex_hull.hpp:76-116
//          constructAndrewMonotoneChainConvexHull              convex_hull.hpp:212-281
//          constructJarvisMarchConvexHull                      convex_hull.hpp:283-335
//          partitionVector / constructChanConvexHull           convex_hull.hpp:337-424
//          findOrderedConvexOutlines / ...ConcaveOutlines      src/polygon_simplification.cpp:31-79, 100-118
// Arithmetic is IEEE float32 with every product and difference rounded on its own (the reference build
// has no FMA contraction: x86-64 baseline, build.sh:13); this file is compiled the same way.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>

namespace
{
struct P2
{
    float x, y;
};

// convex_hull.hpp:51-61: y-major order; equal y = |dy| < epsilon
bool before(const P2 &a, const P2 &b)
{
    return a.y < b.y || (std::fabs(a.y - b.y) < std::numeric_limits<float>::epsilon() && a.x < b.x);
}

// convex_hull.hpp:63-73
bool same(const P2 &a, const P2 &b)
{
    return std::fabs(a.x - b.x) < std::numeric_limits<float>::epsilon() &&
           std::fabs(a.y - b.y) < std::numeric_limits<float>::epsilon();
}

// convex_hull.hpp:76-84; > 0 = counter-clockwise turn o -> a -> b
float turn(const P2 &o, const P2 &a, const P2 &b)
{
    const float ax = a.x - o.x, ay = a.y - o.y, bx = b.x - o.x, by = b.y - o.y;
    const volatile float l = ax * by, r = bx * ay; // two rounded products, then one rounded difference
    return l - r;
}

// convex_hull.hpp:212-281 with COUNTERCLOCKWISE, OPEN: indices into pts of the hull vertices
std::vector<int> monotone_chain(const std::vector<P2> &pts)
{
    const int n = static_cast<int>(pts.size());
    if (n < 3)
        return {};
    std::vector<P2> s(pts);
    std::sort(s.begin(), s.end(), before);
    std::vector<int> st(2 * static_cast<std::size_t>(n));
    int k = 0;
    for (int i = 0; i < n; ++i) // lower chain
    {
        while (k >= 2 && !(turn(s[st[k - 2]], s[st[k - 1]], s[i]) > 0.0f))
            --k;
        st[k++] = i;
    }
    const int floor_k = k + 1;
    for (int i = n - 2; i >= 0; --i) // upper chain
    {
        while (k >= floor_k && !(turn(s[st[k - 2]], s[st[k - 1]], s[i]) > 0.0f))
            --k;
        st[k++] = i;
    }
    st.resize(k - 1);
    for (int &h : st) // first original point equal to the sorted one (stays the sorted position if none)
        for (int j = 0; j < n; ++j)
            if (same(s[h], pts[j]))
            {
                h = j;
                break;
            }
    return st;
}

// convex_hull.hpp:283-335 with COUNTERCLOCKWISE, OPEN
std::vector<int> jarvis(const std::vector<P2> &pts)
{
    const int n = static_cast<int>(pts.size());
    if (n < 3)
        return {};
    int left = 0;
    for (int i = 1; i < n; ++i)
        if (pts[i].x < pts[left].x)
            left = i;
    std::vector<int> hull;
    int p = left;
    do
    {
        hull.push_back(p);
        int q = (p + 1) % n;
        for (int i = 0; i < n; ++i)
            if (turn(pts[p], pts[i], pts[q]) > 0.0f)
                q = i;
        p = q;
    } while (p != left && static_cast<int>(hull.size()) <= n); // the guard is ours: the reference would spin
    return hull;
}

// convex_hull.hpp:366-424
std::vector<int> chan(const std::vector<P2> &pts)
{
    const int n = static_cast<int>(pts.size());
    if (n < 3)
        return {};
    const int subsets = static_cast<int>(std::ceil(std::sqrt(n)));
    const int per = n / subsets, extra = n % subsets;
    std::vector<P2> merged;
    std::vector<int> merged_idx;
    int at = 0;
    for (int s = 0; s < subsets; ++s) // partitionVector: the first `extra` subsets hold one more
    {
        const int size = per + (s < extra ? 1 : 0);
        const std::vector<P2> part(pts.begin() + at, pts.begin() + at + size);
        for (const int h : monotone_chain(part))
        {
            merged.push_back(pts[at + h]);
            merged_idx.push_back(at + h);
        }
        at += size;
    }
    std::vector<int> out;
    const std::vector<int> march = jarvis(merged);
    if (march.size() > merged.size())
        return {-1}; // the march never came back to its start: the reference does not terminate on this input
    for (const int h : march)
        out.push_back(merged_idx[h]);
    return out;
}
} // namespace

extern "C"
{

// clusters as CSR over `points` (stride_floats floats per point, xyz first). mode 0 = findOrderedConvexOutlines
// (monotone chain up to 1000 points, CHAN above), 1 = the convex branch of findOrderedConcaveOutlines (below 20
// points; larger clusters get size 0 here — their concave hull is outside this restatement).
// sizes_out[k] = vertices of cluster k (0xFFFFFFFF: the reference's Jarvis march would not terminate);
// idx_out = cluster-local indices of the vertices, end to end.
long long oracle_convex_outlines(const float *points, const std::uint32_t *offsets, std::uint32_t n_clusters,
                                 std::uint32_t stride_floats, int mode, std::uint32_t *sizes_out,
                                 std::uint32_t *idx_out, long long capacity)
{
    long long total = 0;
    std::vector<P2> pts;
    for (std::uint32_t k = 0; k < n_clusters; ++k)
    {
        pts.clear();
        for (std::uint32_t i = offsets[k]; i < offsets[k + 1]; ++i)
        {
            const float *p = points + static_cast<std::size_t>(i) * stride_floats;
            pts.push_back(P2{p[0], p[1]});
        }
        std::vector<int> hull;
        if (mode == 0)
            hull = pts.size() > 1000 ? chan(pts) : monotone_chain(pts);
        else if (pts.size() < 20)
            hull = monotone_chain(pts);
        if (hull.size() == 1 && hull[0] < 0)
        {
            sizes_out[k] = 0xFFFFFFFFu; // reference would spin forever (Jarvis march that does not close)
            continue;
        }
        sizes_out[k] = static_cast<std::uint32_t>(hull.size());
        if (total + static_cast<long long>(hull.size()) > capacity)
            return -1;
        for (const int h : hull)
            idx_out[total++] = static_cast<std::uint32_t>(h);
    }
    return total;
}

} // extern "C"
