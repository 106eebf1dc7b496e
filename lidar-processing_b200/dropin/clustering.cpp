// Host shell of lidar_processing::Clusterer over the lidar_b200 C ABI.
#include "clustering.hpp"

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

Clusterer::Clusterer()
{
    const int status = lidar_b200_create(0, 200'000U, 1U, &context_);
    if (status != LIDAR_B200_OK)
        throw std::runtime_error("lidar_b200_create failed: no usable CUDA device (there is no CPU fallback)");
    reserve_memory();
}

Clusterer::~Clusterer()
{
    lidar_b200_destroy(context_);
}

void Clusterer::update_configuration(const ClusteringConfiguration &configuration)
{
    lidar_b200_clu_cfg cfg;
    cfg.distance_squared = configuration.distance_squared;
    cfg.cluster_quality = configuration.cluster_quality;
    cfg.min_cluster_size = configuration.min_cluster_size;
    cfg.max_cluster_size = configuration.max_cluster_size;
    const int status = lidar_b200_clu_configure(context_, &cfg);
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::update_configuration");
    configuration_ = configuration;
}

void Clusterer::reserve_memory(std::uint32_t number_of_points)
{
    const int status = lidar_b200_reserve(context_, number_of_points, 1U);
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::reserve_memory");
}

template <typename PointT>
void Clusterer::cluster(const pcl::PointCloud<PointT> &cloud_in, std::vector<ClusteringLabel> &labels)
{
    static_assert(sizeof(PointT) % 4 == 0 && sizeof(PointT) >= 12, "point layout");
    labels.assign(cloud_in.size(), UNDEFINED);
    if (cloud_in.empty())
    {
        return;
    }
    const int status = lidar_b200_cluster(context_, cloud_in.points.data(), static_cast<std::uint32_t>(cloud_in.size()),
                                          static_cast<std::uint32_t>(sizeof(PointT)), labels.data());
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::cluster");
}

template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZ> &cloud_in, std::vector<ClusteringLabel> &labels);

template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZI> &cloud_in, std::vector<ClusteringLabel> &labels);

template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZL> &cloud_in, std::vector<ClusteringLabel> &labels);

template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZRGB> &cloud_in,
                                 std::vector<ClusteringLabel> &labels);

template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZRGBL> &cloud_in,
                                 std::vector<ClusteringLabel> &labels);

} // namespace lidar_processing
