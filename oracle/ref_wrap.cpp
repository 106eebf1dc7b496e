// TEST INFRASTRUCTURE ONLY — C wrapper around the UNMODIFIED reference Clusterer / KDTree.
// The reference sources are compiled where they lie (/root/reference/src, read-only); nothing is
// copied. Built by oracle/Makefile into oracle/_ref/libref_cluster.so (git-ignored, travels to the
// GPU box). Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it.
//
// Wraps: lidar_processing::Clusterer::cluster   (reference src/clustering.cpp:47-125)
//        lidar_processing::KDTree<float,3>       (reference src/kdtree.hpp:174-225, 292-341)
#include "clustering.hpp" // from /root/reference/src
#include "oracle.h"     // restated Segmenter (segmentation.cpp needs Eigen + PCL, absent here)

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

using lidar_processing::Clusterer;
using lidar_processing::ClusteringConfiguration;
using lidar_processing::ClusteringLabel;

extern "C"
{

// points: m records of `stride_floats` floats, xyz first. Returns 0.
int ref_cluster(const float *points, std::uint32_t m, std::uint32_t stride_floats, float distance_squared,
                float cluster_quality, std::uint32_t min_cluster_size, std::uint32_t max_cluster_size,
                std::int32_t *labels_out)
{
    pcl::PointCloud<pcl::PointXYZRGBL> cloud; // the type the processor node uses (processor.cpp:158-163)
    cloud.reserve(m);
    for (std::uint32_t i = 0; i < m; ++i)
    {
        const float *p = points + static_cast<std::size_t>(i) * stride_floats;
        cloud.emplace_back(p[0], p[1], p[2], 0, 255, 0, 1);
    }
    Clusterer clusterer;
    ClusteringConfiguration cfg;
    cfg.distance_squared = distance_squared;
    cfg.cluster_quality = cluster_quality;
    cfg.min_cluster_size = min_cluster_size;
    cfg.max_cluster_size = max_cluster_size;
    clusterer.update_configuration(cfg);
    std::vector<ClusteringLabel> labels;
    clusterer.cluster(cloud, labels);
    if (m)
        std::memcpy(labels_out, labels.data(), sizeof(std::int32_t) * m);
    return 0;
}

// Times `repeats` calls of Clusterer::cluster on a long-lived instance (as the node holds it,
// processor.cpp:132); returns best and mean milliseconds.
int ref_cluster_timed(const float *points, std::uint32_t m, std::uint32_t stride_floats, std::uint32_t repeats,
                      std::int32_t *labels_out, double *best_ms, double *mean_ms)
{
    pcl::PointCloud<pcl::PointXYZRGBL> cloud;
    cloud.reserve(m);
    for (std::uint32_t i = 0; i < m; ++i)
    {
        const float *p = points + static_cast<std::size_t>(i) * stride_floats;
        cloud.emplace_back(p[0], p[1], p[2], 0, 255, 0, 1);
    }
    Clusterer clusterer;
    std::vector<ClusteringLabel> labels;
    double best = 1e300, sum = 0.0;
    for (std::uint32_t r = 0; r < repeats; ++r)
    {
        const auto t0 = std::chrono::steady_clock::now();
        clusterer.cluster(cloud, labels);
        const auto t1 = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        best = ms < best ? ms : best;
        sum += ms;
    }
    if (m && labels_out)
        std::memcpy(labels_out, labels.data(), sizeof(std::int32_t) * m);
    *best_ms = best;
    *mean_ms = repeats ? sum / repeats : 0.0;
    return 0;
}

// Pre-order rank of every point in the reference k-d tree: one radius_search with a huge radius
// prunes nothing and therefore returns all nodes in traversal (= pre-) order (kdtree.hpp:292-341).
int ref_kd_preorder(const float *points, std::uint32_t m, std::uint32_t stride_floats, std::uint32_t *order_out)
{
    using namespace lidar_processing;
    containers::Vector<Point<float, 3>> pts;
    pts.reserve(m);
    for (std::uint32_t i = 0; i < m; ++i)
    {
        const float *p = points + static_cast<std::size_t>(i) * stride_floats;
        pts.push_back({p[0], p[1], p[2]});
    }
    KDTree<float, 3> tree;
    tree.reserve(m);
    tree.rebuild(pts);
    containers::Vector<KDTree<float, 3>::RetT> neigh;
    neigh.reserve(m);
    tree.radius_search(pts[0], 3.0e38F, neigh);
    if (neigh.size() != m)
        return 1;
    for (std::uint32_t i = 0; i < m; ++i)
        order_out[i] = neigh[i].first; // order_out[rank] = point index
    return 0;
}

// radius_search of the reference tree for a list of query point indices; results (index, d2) are
// concatenated, offsets_out has nq+1 entries. Used to pin neighbour sets / d2 bits / order.
int ref_radius_search(const float *points, std::uint32_t m, std::uint32_t stride_floats, const std::uint32_t *queries,
                      std::uint32_t nq, float radius_sqr, std::uint32_t *offsets_out, std::uint32_t *idx_out,
                      float *d2_out, std::uint32_t capacity)
{
    using namespace lidar_processing;
    containers::Vector<Point<float, 3>> pts;
    pts.reserve(m);
    for (std::uint32_t i = 0; i < m; ++i)
    {
        const float *p = points + static_cast<std::size_t>(i) * stride_floats;
        pts.push_back({p[0], p[1], p[2]});
    }
    KDTree<float, 3> tree;
    tree.reserve(m);
    tree.rebuild(pts);
    containers::Vector<KDTree<float, 3>::RetT> neigh;
    std::uint32_t off = 0;
    offsets_out[0] = 0;
    for (std::uint32_t q = 0; q < nq; ++q)
    {
        tree.radius_search(pts[queries[q]], radius_sqr, neigh);
        for (const auto &[k, d] : neigh)
        {
            if (off >= capacity)
                return 2;
            idx_out[off] = k;
            d2_out[off] = d;
            ++off;
        }
        offsets_out[q + 1] = off;
    }
    return 0;
}

// CPU baseline for the whole hot path, the way BASELINE.md §3 asks for it: restated Segmenter
// (std::sort tie order, i.e. what the reference compiles to in this container) followed by the
// UNMODIFIED reference Clusterer on the resulting obstacle cloud (processor.cpp:150-178). One
// long-lived Clusterer per thread (processor.cpp:129-132), frames dealt round-robin to `nthreads`
// std::threads. per_frame_ms[f] = segment + cluster time of frame f on its thread.
int ref_pipeline_run(const float *const *frames, const std::uint32_t *counts, std::uint32_t nframes,
                     std::uint32_t stride_floats, std::uint32_t nthreads, double *per_frame_ms, double *wall_s,
                     std::uint32_t *n_obstacle_out, std::uint32_t *n_clusters_out)
{
    nthreads = nthreads ? nthreads : 1U;
    oracle_seg_cfg seg_cfg;
    oracle_seg_cfg_default(&seg_cfg);
    auto worker = [&](std::uint32_t tid) {
        Clusterer clusterer;
        std::vector<ClusteringLabel> cluster_labels;
        std::vector<std::uint32_t> seg_labels, ground_idx, obstacle_idx;
        pcl::PointCloud<pcl::PointXYZRGBL> obstacle_cloud;
        for (std::uint32_t f = tid; f < nframes; f += nthreads)
        {
            const std::uint32_t n = counts[f];
            const float *pts = frames[f];
            const auto t0 = std::chrono::steady_clock::now();
            seg_labels.assign(n, 0U);
            ground_idx.resize(n ? n : 1U);
            obstacle_idx.resize(n ? n : 1U);
            std::uint32_t ng = 0, no = 0;
            oracle_segment(pts, n, stride_floats, &seg_cfg, 0, seg_labels.data(), ground_idx.data(), &ng,
                           obstacle_idx.data(), &no, nullptr, nullptr);
            obstacle_cloud.clear();
            obstacle_cloud.reserve(no);
            for (std::uint32_t k = 0; k < no; ++k)
            {
                const float *p = pts + static_cast<std::size_t>(obstacle_idx[k]) * stride_floats;
                obstacle_cloud.emplace_back(p[0], p[1], p[2], 0, 255, 0, 1);
            }
            clusterer.cluster(obstacle_cloud, cluster_labels);
            const auto t1 = std::chrono::steady_clock::now();
            per_frame_ms[f] = std::chrono::duration<double, std::milli>(t1 - t0).count();
            if (n_obstacle_out)
                n_obstacle_out[f] = no;
            if (n_clusters_out)
            {
                std::int32_t mx = -1;
                for (const auto l : cluster_labels)
                    mx = l > mx ? l : mx;
                n_clusters_out[f] = static_cast<std::uint32_t>(mx + 1);
            }
        }
    };
    const auto w0 = std::chrono::steady_clock::now();
    std::vector<std::thread> threads;
    for (std::uint32_t t = 1; t < nthreads; ++t)
        threads.emplace_back(worker, t);
    worker(0U);
    for (auto &t : threads)
        t.join();
    *wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
    return 0;
}

// The same pipeline over `n_passes` passes of the whole frame list with LONG-LIVED workers: every thread constructs its
// Clusterer and buffers once (as the node does, processor.cpp:129-132), all threads meet at a barrier before and after
// every pass, pass_wall_s[p] is the time between the two barriers of pass p. bench.py uses the first passes as
// warm-up and the rest as its timed steps, so nothing is constructed inside a timed region.
int ref_pipeline_passes(const float *const *frames, const std::uint32_t *counts, std::uint32_t nframes,
                        std::uint32_t stride_floats, std::uint32_t nthreads, std::uint32_t n_passes, double *pass_wall_s,
                        double *per_frame_ms /* last pass */, std::uint32_t *n_obstacle_out, std::uint32_t *n_clusters_out)
{
    nthreads = nthreads ? nthreads : 1U;
    oracle_seg_cfg seg_cfg;
    oracle_seg_cfg_default(&seg_cfg);
    std::mutex mu;
    std::condition_variable cv;
    std::uint32_t arrived = 0, generation = 0;
    auto barrier = [&]() {
        std::unique_lock<std::mutex> lk(mu);
        const std::uint32_t gen = generation;
        if (++arrived == nthreads)
        {
            arrived = 0;
            ++generation;
            cv.notify_all();
        }
        else
            cv.wait(lk, [&] { return generation != gen; });
    };
    auto worker = [&](std::uint32_t tid) {
        Clusterer clusterer;
        std::vector<ClusteringLabel> cluster_labels;
        std::vector<std::uint32_t> seg_labels, ground_idx, obstacle_idx;
        pcl::PointCloud<pcl::PointXYZRGBL> obstacle_cloud;
        for (std::uint32_t pass = 0; pass < n_passes; ++pass)
        {
            barrier();
            const auto p0 = std::chrono::steady_clock::now();
            for (std::uint32_t f = tid; f < nframes; f += nthreads)
            {
                const std::uint32_t n = counts[f];
                const float *pts = frames[f];
                const auto t0 = std::chrono::steady_clock::now();
                seg_labels.assign(n, 0U);
                ground_idx.resize(n ? n : 1U);
                obstacle_idx.resize(n ? n : 1U);
                std::uint32_t ng = 0, no = 0;
                oracle_segment(pts, n, stride_floats, &seg_cfg, 0, seg_labels.data(), ground_idx.data(), &ng,
                               obstacle_idx.data(), &no, nullptr, nullptr);
                obstacle_cloud.clear();
                obstacle_cloud.reserve(no);
                for (std::uint32_t k = 0; k < no; ++k)
                {
                    const float *p = pts + static_cast<std::size_t>(obstacle_idx[k]) * stride_floats;
                    obstacle_cloud.emplace_back(p[0], p[1], p[2], 0, 255, 0, 1);
                }
                clusterer.cluster(obstacle_cloud, cluster_labels);
                const auto t1 = std::chrono::steady_clock::now();
                per_frame_ms[f] = std::chrono::duration<double, std::milli>(t1 - t0).count();
                if (n_obstacle_out)
                    n_obstacle_out[f] = no;
                if (n_clusters_out)
                {
                    std::int32_t mx = -1;
                    for (const auto l : cluster_labels)
                        mx = l > mx ? l : mx;
                    n_clusters_out[f] = static_cast<std::uint32_t>(mx + 1);
                }
            }
            barrier();
            if (tid == 0U)
                pass_wall_s[pass] = std::chrono::duration<double>(std::chrono::steady_clock::now() - p0).count();
        }
    };
    std::vector<std::thread> threads;
    for (std::uint32_t t = 1; t < nthreads; ++t)
        threads.emplace_back(worker, t);
    worker(0U);
    for (auto &t : threads)
        t.join();
    return 0;
}

} // extern "C"
