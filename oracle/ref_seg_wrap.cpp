// TEST INFRASTRUCTURE ONLY — C wrapper around the UNMODIFIED reference Segmenter.
// /root/reference/src/segmentation.cpp is compiled where it lies against two stand-ins: oracle/pcl_shim (point
// structs + PointCloud container) and oracle/eigen_shim (the Eigen API subset the file uses, with this repo's
// restated arithmetic — see the header of oracle/eigen_shim/Eigen/Dense for what that does and does not pin).
// Built by oracle/Makefile into oracle/_ref/libref_segment.so (git-ignored, travels to the GPU box).
//
// Wraps: lidar_processing::Segmenter::segment<pcl::PointXYZI>   (reference src/segmentation.cpp:311-345)
#include "segmentation.hpp" // from /root/reference/src
#include "oracle.h"

#include <cstdint>
#include <cstring>
#include <vector>

using lidar_processing::SegmentationConfiguration;
using lidar_processing::SegmentationLabel;
using lidar_processing::Segmenter;

namespace
{
// The output clouds carry whole point copies, not indices (segmentation.cpp:335,342): the original index rides in
// the intensity field (exact in float32 up to 2^24 points; the hot path never reads intensity).
void fill_cloud(const float *pts, std::uint32_t n, std::uint32_t stride, pcl::PointCloud<pcl::PointXYZI> &cloud)
{
    cloud.clear();
    cloud.reserve(n);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        const float *p = pts + static_cast<std::size_t>(i) * stride;
        cloud.emplace_back(p[0], p[1], p[2], static_cast<float>(i));
    }
}

SegmentationConfiguration to_cfg(const oracle_seg_cfg *cfg)
{
    SegmentationConfiguration c;
    c.sensor_height_m = cfg->sensor_height_m;
    c.orthogonal_distance_threshold = cfg->orthogonal_distance_threshold;
    c.initial_seed_threshold = cfg->initial_seed_threshold;
    c.number_of_iterations = cfg->number_of_iterations;
    c.number_of_planar_partitions = cfg->number_of_planar_partitions;
    c.number_of_lower_point_representatives = cfg->number_of_lower_point_representatives;
    return c;
}
} // namespace

extern "C"
{

// Same contract as oracle_segment (oracle.h) with tie_mode 0 (the std::sort this toolchain compiles the reference's
// std::sort(std::execution::par, ...) to). labels: n entries IN/OUT — they are the caller's vector as the reference's
// labels.resize() leaves it (segmentation.cpp:315). Returns 0; -2 when n >= 2^24.
int ref_segment(const float *pts, std::uint32_t n, std::uint32_t stride_floats, const oracle_seg_cfg *cfg,
                std::uint32_t *labels, std::uint32_t *ground_idx, std::uint32_t *n_ground, std::uint32_t *obstacle_idx,
                std::uint32_t *n_obstacle)
{
    if (n >= (1U << 24))
        return -2;
    Segmenter segmenter;
    segmenter.update_configuration(to_cfg(cfg));
    pcl::PointCloud<pcl::PointXYZI> cloud, ground, obstacle;
    fill_cloud(pts, n, stride_floats, cloud);
    std::vector<SegmentationLabel> lab(n);
    for (std::uint32_t i = 0; i < n; ++i)
        lab[i] = static_cast<SegmentationLabel>(labels[i]);
    segmenter.segment(cloud, lab, ground, obstacle);
    for (std::uint32_t i = 0; i < n; ++i)
        labels[i] = static_cast<std::uint32_t>(lab[i]);
    *n_ground = static_cast<std::uint32_t>(ground.size());
    *n_obstacle = static_cast<std::uint32_t>(obstacle.size());
    for (std::size_t i = 0; i < ground.size(); ++i)
        ground_idx[i] = static_cast<std::uint32_t>(ground[i].intensity);
    for (std::size_t i = 0; i < obstacle.size(); ++i)
        obstacle_idx[i] = static_cast<std::uint32_t>(obstacle[i].intensity);
    return 0;
}

// Two consecutive frames through ONE long-lived Segmenter and ONE labels vector, the way the node holds them
// (processor.cpp:129-131,150): the second call's labels show the stale entries of the first (segmentation.cpp:315).
int ref_segment_pair(const float *pts_a, std::uint32_t n_a, const float *pts_b, std::uint32_t n_b,
                     std::uint32_t stride_floats, const oracle_seg_cfg *cfg, std::uint32_t *labels_b_out)
{
    if (n_a >= (1U << 24) || n_b >= (1U << 24))
        return -2;
    Segmenter segmenter;
    segmenter.update_configuration(to_cfg(cfg));
    pcl::PointCloud<pcl::PointXYZI> cloud, ground, obstacle;
    std::vector<SegmentationLabel> lab;
    fill_cloud(pts_a, n_a, stride_floats, cloud);
    segmenter.segment(cloud, lab, ground, obstacle);
    fill_cloud(pts_b, n_b, stride_floats, cloud);
    segmenter.segment(cloud, lab, ground, obstacle);
    for (std::size_t i = 0; i < lab.size(); ++i)
        labels_b_out[i] = static_cast<std::uint32_t>(lab[i]);
    return static_cast<int>(lab.size());
}

} // extern "C"
