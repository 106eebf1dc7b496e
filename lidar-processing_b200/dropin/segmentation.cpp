// Host shell of lidar_processing::Segmenter over the lidar_b200 C ABI.
#include "segmentation.hpp"

#include "lidar_b200.h"

#include <stdexcept>
#include <string>

namespace lidar_processing
{
namespace
{
[[noreturn]] void raise(lidar_b200_ctx *context, int status, const char *what)
{
    const std::string message = std::string(what) + ": " + (context ? lidar_b200_last_error(context) : "no CUDA context");
    if (status == LIDAR_B200_ERR_UNSUPPORTED || status == LIDAR_B200_ERR_INVALID)
        throw std::invalid_argument(message);
    throw std::runtime_error(message);
}
} // namespace

Segmenter::Segmenter()
{
    const int status = lidar_b200_create(0, 200'000U, 1U, &context_);
    if (status != LIDAR_B200_OK)
        throw std::runtime_error("lidar_b200_create failed: no usable CUDA device (there is no CPU fallback)");
    reserve_memory();
}

Segmenter::~Segmenter()
{
    lidar_b200_destroy(context_);
}

void Segmenter::update_configuration(const SegmentationConfiguration &configuration)
{
    lidar_b200_seg_cfg cfg;
    cfg.sensor_height_m = configuration.sensor_height_m;
    cfg.orthogonal_distance_threshold = configuration.orthogonal_distance_threshold;
    cfg.initial_seed_threshold = configuration.initial_seed_threshold;
    cfg.number_of_iterations = configuration.number_of_iterations;
    cfg.number_of_planar_partitions = configuration.number_of_planar_partitions;
    cfg.number_of_lower_point_representatives = configuration.number_of_lower_point_representatives;
    const int status = lidar_b200_seg_configure(context_, &cfg);
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Segmenter::update_configuration");
    configuration_ = configuration;
}

void Segmenter::reserve_memory(std::uint32_t number_of_points)
{
    const int status = lidar_b200_reserve(context_, number_of_points, 1U);
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Segmenter::reserve_memory");
    ground_indices_.reserve(number_of_points);
    obstacle_indices_.reserve(number_of_points);
}

template <typename PointT>
void Segmenter::segment(const pcl::PointCloud<PointT> &cloud_in, std::vector<SegmentationLabel> &labels,
                        pcl::PointCloud<PointT> &ground_cloud, pcl::PointCloud<PointT> &obstacle_cloud)
{
    static_assert(sizeof(SegmentationLabel) == sizeof(std::uint32_t), "label layout");
    static_assert(sizeof(PointT) % 4 == 0 && sizeof(PointT) >= 12, "point layout");

    labels.resize(cloud_in.size(), SegmentationLabel::UNKNOWN); // keeps old entries, like the reference
    ground_cloud.clear();
    obstacle_cloud.clear();

    const std::uint32_t number_of_points = static_cast<std::uint32_t>(cloud_in.points.size());
    if (number_of_points == 0U)
    {
        return;
    }

    ground_indices_.resize(number_of_points);
    obstacle_indices_.resize(number_of_points);
    std::uint32_t number_of_ground = 0U;
    std::uint32_t number_of_obstacle = 0U;
    const int status = lidar_b200_segment(context_, cloud_in.points.data(), number_of_points,
                                          static_cast<std::uint32_t>(sizeof(PointT)),
                                          reinterpret_cast<std::uint32_t *>(labels.data()), ground_indices_.data(),
                                          &number_of_ground, obstacle_indices_.data(), &number_of_obstacle);
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Segmenter::segment");

    ground_cloud.reserve(number_of_ground);
    for (std::uint32_t i = 0U; i < number_of_ground; ++i)
    {
        ground_cloud.push_back(cloud_in[ground_indices_[i]]);
    }
    obstacle_cloud.reserve(number_of_obstacle);
    for (std::uint32_t i = 0U; i < number_of_obstacle; ++i)
    {
        obstacle_cloud.push_back(cloud_in[obstacle_indices_[i]]);
    }
}

template void Segmenter::segment(const pcl::PointCloud<pcl::PointXYZ> &cloud_in, std::vector<SegmentationLabel> &labels,
                                 pcl::PointCloud<pcl::PointXYZ> &ground_cloud,
                                 pcl::PointCloud<pcl::PointXYZ> &obstacle_cloud);

template void Segmenter::segment(const pcl::PointCloud<pcl::PointXYZI> &cloud_in,
                                 std::vector<SegmentationLabel> &labels, pcl::PointCloud<pcl::PointXYZI> &ground_cloud,
                                 pcl::PointCloud<pcl::PointXYZI> &obstacle_cloud);

} // namespace lidar_processing
