// Drop-in replacement for the reference's src/clustering.hpp (YevgeniyEngineer/LiDAR-Processing):
// same namespace, label alias, configuration struct, class name, constants, public member functions
// and explicit instantiations (reference src/clustering.hpp:38-91). src/polygonization.hpp includes
// this header for ClusteringLabel (polygonization.hpp:26,71) and keeps compiling. The body runs on a
// B200 through the C ABI in include/lidar_b200.h; there is no CPU implementation behind it.
#ifndef LIDAR_PROCESSING__CLUSTERING_HPP
#define LIDAR_PROCESSING__CLUSTERING_HPP

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <cstdint>
#include <limits>
#include <vector>

struct lidar_b200_ctx;

namespace lidar_processing
{
using ClusteringLabel = std::int32_t;

struct ClusteringConfiguration final
{
    float distance_squared{0.18F};
    float cluster_quality{0.5F};
    std::uint32_t min_cluster_size{4U};
    std::uint32_t max_cluster_size{std::numeric_limits<std::uint32_t>::max()};
};

class Clusterer final
{
  public:
    static constexpr ClusteringLabel UNDEFINED{std::numeric_limits<std::int32_t>::lowest()};
    static constexpr ClusteringLabel INVALID{-1};

    Clusterer();
    ~Clusterer();
    Clusterer(const Clusterer &) = delete;
    Clusterer &operator=(const Clusterer &) = delete;

    void update_configuration(const ClusteringConfiguration &configuration);

    void reserve_memory(std::uint32_t number_of_points = 200'000U);

    // Same contract as the reference (src/clustering.cpp:47-125): labels.assign(n, UNDEFINED), then
    // every point receives a dense cluster id 0..K-1 or INVALID. The partition is bit-identical to
    // the reference's for the same cloud in the same point order. Throws std::runtime_error on a
    // CUDA failure or non-finite coordinates.
    template <typename PointT>
    void cluster(const pcl::PointCloud<PointT> &cloud_in, std::vector<ClusteringLabel> &labels);

    // Extension (not in the reference class): the split of the cloud by cluster label that the reference's
    // caller does on the host right after cluster() (reference src/processor.cpp:180-200), done on the
    // device for the cloud of the LAST cluster() call. clustered_cloud[k] receives PointXYZ(x, y, z) of
    // every point with label k in ascending point index; INVALID points are skipped.
    void split_last_clusters(std::vector<pcl::PointCloud<pcl::PointXYZ>> &clustered_cloud);

    // Extension: ordered outlines (convex: counter-clockwise, open) of the clusters of the LAST
    // split_last_clusters() call, computed on the device bit-identically to the reference's host functions.
    // OutlinePoint has the layout of geom::Point<float> (reference Convex-Hull/convex_hull.hpp:42-49).
    //   CONVEX        = findOrderedConvexOutlines (reference src/polygon_simplification.cpp:31-79): every cluster.
    //   CONCAVE_SMALL = the convex branch of findOrderedConcaveOutlines (:100-118): clusters below 20 points;
    //                   the ids of the larger ones are returned in host_clusters — the caller runs the
    //                   reference's geometry::ConcaveHull on those (it stays on the host) and stores the result
    //                   in outlines[id].
    //   CONCAVE       = findOrderedConcaveOutlines as a whole (:81-149): that branch below 20 points and, from 20
    //                   points on, geometry::ConcaveHull with chi = 0.2 (Concave-Hull/concave_hull.hpp:96-193 over
    //                   delaunator.cpp) re-enacted on the device value for value; those outlines are closed (first
    //                   vertex repeated) like the reference's. host_clusters stays empty. A cluster on which the
    //                   reference throws "not triangulation" (20+ points collinear in x, y) makes this call throw too.
    // outlines[k] belongs to clustered_cloud[k]; an empty outline is one the reference drops from its output.
    struct OutlinePoint
    {
        float x, y;
    };
    enum class OutlinePolicy : std::uint32_t
    {
        CONVEX = 0U,
        CONCAVE_SMALL = 1U,
        CONCAVE = 2U
    };
    void outline_last_clusters(OutlinePolicy policy, std::vector<std::vector<OutlinePoint>> &outlines,
                               std::vector<std::uint32_t> &host_clusters);

    // Extension: convertClusteredCloudToColorizedCloud (reference src/conversions.cpp:32-60) for the clusters of the
    // LAST split_last_clusters() call: one colour per cluster drawn with std::rand() % 256 for r, g, b in this
    // order (the caller's process-wide sequence continues exactly as in the reference), the 32-byte PointXYZRGB
    // records are built on the device.
    void colorize_last_clusters(pcl::PointCloud<pcl::PointXYZRGB> &colorized_cloud);

    // Extension: the point lists of convertPointXYZTypeToMarkerArray (reference src/conversions.hpp:72-120) for the
    // outlines of the LAST outline_last_clusters() call: per non-empty outline its vertices as {double x, y, 0.0}
    // (the layout of geometry_msgs::msg::Point) with the first vertex repeated at the end; marker_ids[i] is the
    // cluster the i-th strip belongs to.
    struct MarkerPoint
    {
        double x, y, z;
    };
    void marker_points_of_last_outlines(std::vector<std::vector<MarkerPoint>> &strips, std::vector<std::uint32_t> &marker_ids);

  private:
    lidar_b200_ctx *context_{nullptr};
    ClusteringConfiguration configuration_{};
    std::uint32_t last_cloud_size_{0U};
    std::vector<std::uint32_t> split_offsets_;
    std::vector<float> split_points_;
    std::vector<std::uint32_t> outline_offsets_;
    std::vector<float> outline_xy_;
    std::uint32_t split_clusters_{0U};
    std::vector<std::uint32_t> cluster_colors_, marker_offsets_;
    std::vector<float> colorized_records_;
    std::vector<double> marker_points_;
};

extern template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZ> &cloud_in,
                                        std::vector<ClusteringLabel> &labels);

extern template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZI> &cloud_in,
                                        std::vector<ClusteringLabel> &labels);

extern template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZL> &cloud_in,
                                        std::vector<ClusteringLabel> &labels);

extern template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZRGB> &cloud_in,
                                        std::vector<ClusteringLabel> &labels);

extern template void Clusterer::cluster(const pcl::PointCloud<pcl::PointXYZRGBL> &cloud_in,
                                        std::vector<ClusteringLabel> &labels);
} // namespace lidar_processing

#endif // LIDAR_PROCESSING__CLUSTERING_HPP
