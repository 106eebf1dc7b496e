// Drop-in replacement for the reference's src/segmentation.hpp (YevgeniyEngineer/LiDAR-Processing):
// same namespace, label enum, configuration struct, class name, public member functions and
// explicit instantiations (reference src/segmentation.hpp:39-141), so src/processor.cpp compiles
// against it unchanged. The body runs on a B200 through the C ABI in include/lidar_b200.h; there is
// no CPU implementation behind it. Private state differs from the reference (which is allowed: the
// reference's own members are private too).
#ifndef LIDAR_PROCESSING__SEGMENTATION_HPP
#define LIDAR_PROCESSING__SEGMENTATION_HPP

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <cstdint>
#include <memory>
#include <vector>

struct lidar_b200_ctx;

namespace lidar_processing
{
enum class SegmentationLabel : std::uint32_t
{
    UNKNOWN = 0U,
    GROUND,
    OBSTACLE
};

struct SegmentationConfiguration final
{
    float sensor_height_m{1.73F};
    float orthogonal_distance_threshold{0.3F};
    float initial_seed_threshold{0.6F};
    std::uint32_t number_of_iterations{3U};
    std::uint32_t number_of_planar_partitions{2U};
    std::uint32_t number_of_lower_point_representatives{5000U};
};

class Segmenter final
{
  public:
    Segmenter();
    ~Segmenter();
    Segmenter(const Segmenter &) = delete;
    Segmenter &operator=(const Segmenter &) = delete;

    // Throws std::invalid_argument for configurations outside the CUDA path's envelope
    // (0 partitions / 0 iterations / 0 representatives); see DESIGN.md.
    void update_configuration(const SegmentationConfiguration &configuration);

    void reserve_memory(std::uint32_t number_of_points = 200'000U);

    // Same contract as the reference (src/segmentation.cpp:311-345): labels is resized (old
    // entries are kept), both clouds are cleared and refilled in x-ascending order per partition.
    // A CUDA failure throws std::runtime_error (the node's main catches std::exception,
    // reference src/processor.cpp:281-285).
    template <typename PointT>
    void segment(const pcl::PointCloud<PointT> &cloud_in, std::vector<SegmentationLabel> &labels,
                 pcl::PointCloud<PointT> &ground_cloud, pcl::PointCloud<PointT> &obstacle_cloud);

  private:
    lidar_b200_ctx *context_{nullptr};
    SegmentationConfiguration configuration_{};
    std::vector<std::uint32_t> ground_indices_;
    std::vector<std::uint32_t> obstacle_indices_;
};

extern template void Segmenter::segment(const pcl::PointCloud<pcl::PointXYZ> &cloud_in,
                                        std::vector<SegmentationLabel> &labels,
                                        pcl::PointCloud<pcl::PointXYZ> &ground_cloud,
                                        pcl::PointCloud<pcl::PointXYZ> &obstacle_cloud);

extern template void Segmenter::segment(const pcl::PointCloud<pcl::PointXYZI> &cloud_in,
                                        std::vector<SegmentationLabel> &labels,
                                        pcl::PointCloud<pcl::PointXYZI> &ground_cloud,
                                        pcl::PointCloud<pcl::PointXYZI> &obstacle_cloud);

} // namespace lidar_processing

#endif // LIDAR_PROCESSING__SEGMENTATION_HPP
