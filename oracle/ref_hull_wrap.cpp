// TEST INFRASTRUCTURE ONLY — C wrapper around the UNMODIFIED reference outline functions.
// The reference sources are compiled where they lie (/root/reference/src/polygon_simplification.cpp,
// /root/reference/Convex-Hull/convex_hull.hpp, /root/reference/Concave-Hull/{concave_hull.hpp,
// delaunator.cpp}); nothing is copied. Built by oracle/Makefile into oracle/_ref/libref_hull.so
// (git-ignored, travels to the GPU box). Only tests/ and bench.py's CPU legs may load it.
//
// Wraps: lidar_processing::findOrderedConvexOutlines   (reference src/polygon_simplification.cpp:31-79)
//        lidar_processing::findOrderedConcaveOutlines  (reference src/polygon_simplification.cpp:81-149)
#include "polygon_simplification.hpp" // from /root/reference/src

#include <cstdint>
#include <exception>
#include <vector>

extern "C"
{

// clusters as CSR: cluster k = points[offsets[k] .. offsets[k+1]) with `stride_floats` floats per point (xyz first).
// mode 0 = findOrderedConvexOutlines, 1 = findOrderedConcaveOutlines. Every cluster is passed on its own so that
// the outline can be attributed to it (the reference drops empty outlines from its output vector): sizes_out[k] =
// vertices of cluster k's outline (0 = dropped, 0xFFFFFFFF = the reference threw), xy_out = the outlines end to end. Returns the number of vertices,
// or -1 when xy_out (capacity in vertices) is too small.
long long ref_outlines(const float *points, const std::uint32_t *offsets, std::uint32_t n_clusters,
                       std::uint32_t stride_floats, int mode, std::uint32_t *sizes_out, float *xy_out,
                       long long capacity)
{
    long long total = 0;
    std::vector<pcl::PointCloud<pcl::PointXYZ>> one(1);
    std::vector<std::vector<geom::Point<float>>> outlines;
    for (std::uint32_t k = 0; k < n_clusters; ++k)
    {
        one[0].clear();
        for (std::uint32_t i = offsets[k]; i < offsets[k + 1]; ++i)
        {
            const float *p = points + static_cast<std::size_t>(i) * stride_floats;
            one[0].emplace_back(p[0], p[1], p[2]);
        }
        try
        {
            if (mode == 0)
                lidar_processing::findOrderedConvexOutlines(one, outlines);
            else
                lidar_processing::findOrderedConcaveOutlines(one, outlines);
        }
        catch (const std::exception &)
        {
            // delaunator throws on degenerate input (e.g. >= 20 collinear points); the node would go down with it
            sizes_out[k] = 0xFFFFFFFFu;
            continue;
        }
        sizes_out[k] = outlines.empty() ? 0u : static_cast<std::uint32_t>(outlines[0].size());
        if (!outlines.empty())
        {
            if (total + static_cast<long long>(outlines[0].size()) > capacity)
                return -1;
            for (const auto &v : outlines[0])
            {
                xy_out[2 * total] = v.x;
                xy_out[2 * total + 1] = v.y;
                ++total;
            }
        }
    }
    return total;
}

} // extern "C"
