// Host shell of lidar_processing::Clusterer over the lidar_b200 C ABI.
#include "clustering.hpp"

#include "lidar_b200.h"

#include <cstdlib>
#include <cstring>
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
    last_cloud_size_ = static_cast<std::uint32_t>(cloud_in.size());
    if (cloud_in.empty())
    {
        return;
    }
    const int status = lidar_b200_cluster(context_, cloud_in.points.data(), static_cast<std::uint32_t>(cloud_in.size()),
                                          static_cast<std::uint32_t>(sizeof(PointT)), labels.data());
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::cluster");
}

void Clusterer::split_last_clusters(std::vector<pcl::PointCloud<pcl::PointXYZ>> &clustered_cloud)
{
    static_assert(sizeof(pcl::PointXYZ) == 16, "the device writes 16-byte PointXYZ records");
    clustered_cloud.clear();
    if (last_cloud_size_ == 0U)
    {
        return;
    }
    int status = lidar_b200_batch_group_clusters(context_);
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::split_last_clusters");
    const std::size_t padded = (static_cast<std::size_t>(last_cloud_size_) + 31U) & ~static_cast<std::size_t>(31U);
    split_offsets_.assign(padded + 1U, 0U);
    split_points_.resize(padded * 4U);
    std::uint32_t number_of_clusters = 0U;
    status = lidar_b200_batch_fetch_clusters(context_, &number_of_clusters, split_offsets_.data(), split_points_.data(), nullptr);
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::split_last_clusters");
    clustered_cloud.resize(number_of_clusters);
    split_clusters_ = number_of_clusters;
    const auto *records = reinterpret_cast<const pcl::PointXYZ *>(split_points_.data());
    for (std::uint32_t k = 0U; k < number_of_clusters; ++k)
    {
        clustered_cloud[k].points.assign(records + split_offsets_[k], records + split_offsets_[k + 1U]);
        clustered_cloud[k].width = static_cast<std::uint32_t>(clustered_cloud[k].points.size());
        clustered_cloud[k].height = 1U;
    }
}

void Clusterer::outline_last_clusters(OutlinePolicy policy, std::vector<std::vector<OutlinePoint>> &outlines,
                                      std::vector<std::uint32_t> &host_clusters)
{
    static_assert(sizeof(OutlinePoint) == 8, "the device writes 8-byte (x, y) records");
    outlines.clear();
    host_clusters.clear();
    if (last_cloud_size_ == 0U)
    {
        return;
    }
    int status = lidar_b200_batch_hull_outlines(context_, static_cast<std::uint32_t>(policy));
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::outline_last_clusters");
    const std::size_t padded = (static_cast<std::size_t>(last_cloud_size_) + 31U) & ~static_cast<std::size_t>(31U);
    outline_offsets_.assign(padded + 1U, 0U);
    outline_xy_.resize(padded * 2U);
    std::uint32_t number_of_vertices = 0U;
    status = lidar_b200_batch_fetch_hulls(context_, &number_of_vertices, outline_offsets_.data(), outline_xy_.data(), nullptr);
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::outline_last_clusters");
    // split_offsets_ still holds the CSR of the split these outlines belong to
    std::uint32_t number_of_clusters = 0U;
    status = lidar_b200_batch_fetch_clusters(context_, &number_of_clusters, nullptr, nullptr, nullptr);
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::outline_last_clusters");
    outlines.resize(number_of_clusters);
    const auto *records = reinterpret_cast<const OutlinePoint *>(outline_xy_.data());
    for (std::uint32_t k = 0U; k < number_of_clusters; ++k)
    {
        outlines[k].assign(records + outline_offsets_[k], records + outline_offsets_[k + 1U]);
        if (policy == OutlinePolicy::CONCAVE_SMALL && split_offsets_[k + 1U] - split_offsets_[k] >= 20U)
            host_clusters.push_back(k); // reference src/polygon_simplification.cpp:100, 119-140
    }
}

void Clusterer::colorize_last_clusters(pcl::PointCloud<pcl::PointXYZRGB> &colorized_cloud)
{
    static_assert(sizeof(pcl::PointXYZRGB) == 32, "the device writes 32-byte PointXYZRGB records");
    colorized_cloud.clear();
    if (last_cloud_size_ == 0U)
    {
        return;
    }
    cluster_colors_.resize(split_clusters_);
    for (auto &word : cluster_colors_)
    {
        // conversions.cpp:49-51: r, g, b in this order, one draw each
        const auto r = static_cast<std::uint32_t>(std::rand() % 256);
        const auto g = static_cast<std::uint32_t>(std::rand() % 256);
        const auto b = static_cast<std::uint32_t>(std::rand() % 256);
        word = (r << 16U) | (g << 8U) | b;
    }
    const std::size_t padded = (static_cast<std::size_t>(last_cloud_size_) + 31U) & ~static_cast<std::size_t>(31U);
    colorized_records_.resize(padded * 8U);
    const int status = lidar_b200_batch_fetch_colorized(context_, cluster_colors_.data(), cluster_colors_.size(),
                                                        colorized_records_.data());
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::colorize_last_clusters");
    const std::size_t number_of_points = split_clusters_ ? split_offsets_[split_clusters_] : 0U;
    colorized_cloud.points.resize(number_of_points);
    if (number_of_points != 0U)
        std::memcpy(static_cast<void *>(colorized_cloud.points.data()), colorized_records_.data(), number_of_points * 32U);
    colorized_cloud.width = static_cast<std::uint32_t>(number_of_points);
    colorized_cloud.height = 1U;
}

void Clusterer::marker_points_of_last_outlines(std::vector<std::vector<MarkerPoint>> &strips,
                                               std::vector<std::uint32_t> &marker_ids)
{
    static_assert(sizeof(MarkerPoint) == 24, "the device writes three doubles per marker point");
    strips.clear();
    marker_ids.clear();
    if (last_cloud_size_ == 0U)
    {
        return;
    }
    const std::size_t padded = (static_cast<std::size_t>(last_cloud_size_) + 31U) & ~static_cast<std::size_t>(31U);
    marker_offsets_.assign(padded + 1U, 0U);
    marker_points_.resize(padded * 6U);
    std::uint32_t number_of_markers = 0U;
    const int status = lidar_b200_batch_fetch_marker_points(context_, &number_of_markers, marker_offsets_.data(), marker_points_.data());
    if (status != LIDAR_B200_OK)
        raise(context_, status, "Clusterer::marker_points_of_last_outlines");
    const auto *records = reinterpret_cast<const MarkerPoint *>(marker_points_.data());
    strips.reserve(number_of_markers);
    for (std::uint32_t k = 0U; k < split_clusters_; ++k)
    {
        // outline k owns marker points [outline_offset[k] + markers_before[k], outline_offset[k+1] + markers_before[k+1])
        const std::size_t first = static_cast<std::size_t>(outline_offsets_[k]) + marker_offsets_[k];
        const std::size_t last = static_cast<std::size_t>(outline_offsets_[k + 1U]) + marker_offsets_[k + 1U];
        if (last > first)
        {
            strips.emplace_back(records + first, records + last);
            marker_ids.push_back(k);
        }
    }
}

template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZ> &cloud_in, std::vector<ClusteringLabel> &labels);

template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZI> &cloud_in, std::vector<ClusteringLabel> &labels);

template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZL> &cloud_in, std::vector<ClusteringLabel> &labels);

template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZRGB> &cloud_in,
                                 std::vector<ClusteringLabel> &labels);

template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZRGBL> &cloud_in,
                                 std::vector<ClusteringLabel> &labels);

} // namespace lidar_processing
