// TEST INFRASTRUCTURE ONLY — the UNMODIFIED reference node (src/processor.cpp) run in-process.
// processor.cpp, conversions.cpp, segmentation.cpp, clustering.cpp and polygon_simplification.cpp are compiled where they
// lie against stand-ins for the libraries this image lacks: oracle/ros_shim (rclcpp + message structs: a subscription is
// a stored callback, a publisher a per-topic sink), oracle/pcl_shim (point structs, PointCloud) and oracle/eigen_shim (see
// its header for what the Eigen stand-in does and does not pin). The node's own `main` is renamed by the build
// (-Dmain=ref_processor_main); rclcpp::spin() of the stand-in calls the hook installed here, which delivers one
// sensor_msgs::msg::PointCloud2 per frame to the node's "pointcloud" callback (Processor::process, processor.cpp:135)
// exactly as the dataloader node builds it (dataloader.cpp:102-139: 32-byte point_step, x y z at 0 4 8, intensity at 16)
// and records what the node publishes. Built by oracle/Makefile into oracle/_ref/libref_node.so.
//
// This is the caller of the hot path: what it does between and after Segmenter::segment / Clusterer::cluster is what
// SURVEY §8(f) rows 1 (per-cluster split, processor.cpp:180-200), 3 (outlines, :212-214) and 4 (colourised cloud and
// MarkerArray, :248-267 with conversions.cpp:32-60, 88-113 and conversions.hpp:72-120) restate on the device.
#include <rclcpp/rclcpp.hpp>
#include <sensor_msgs/msg/point_cloud2.hpp>
#include <visualization_msgs/msg/marker_array.hpp>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

int ref_processor_main(int argc, const char **const argv); // src/processor.cpp's main, renamed by the build

namespace
{
using PointCloud2 = sensor_msgs::msg::PointCloud2;
using MarkerArray = visualization_msgs::msg::MarkerArray;

struct Captured
{
    std::vector<std::uint8_t> ground, obstacle, clustered; // PointCloud2 payloads
    std::uint32_t ground_step{0}, obstacle_step{0}, clustered_step{0};
    std::vector<std::uint32_t> marker_sizes;                // points per marker (closed strip: outline + 1)
    std::vector<std::int32_t> marker_ids;
    std::vector<double> marker_xyz;                         // all marker points end to end
    bool has_ground{false}, has_obstacle{false}, has_clustered{false}, has_markers{false};
};

Captured g_out;
const float *g_points = nullptr;
std::uint32_t g_n = 0, g_stride = 0;
std::string g_error;

PointCloud2 make_message(const float *pts, std::uint32_t n, std::uint32_t stride)
{
    PointCloud2 msg; // dataloader.cpp:102-139
    msg.header.frame_id = "pointcloud";
    msg.header.stamp.sec = 1;
    msg.height = 1;
    msg.width = n;
    msg.is_bigendian = false;
    msg.is_dense = true;
    msg.point_step = 32;
    msg.row_step = 32 * n;
    const char *names[4] = {"x", "y", "z", "intensity"};
    const std::uint32_t offsets[4] = {0, 4, 8, 16};
    for (int i = 0; i < 4; ++i)
    {
        sensor_msgs::msg::PointField f;
        f.name = names[i];
        f.offset = offsets[i];
        f.datatype = sensor_msgs::msg::PointField::FLOAT32;
        f.count = 1;
        msg.fields.push_back(f);
    }
    msg.data.assign(static_cast<std::size_t>(n) * 32, 0);
    for (std::uint32_t i = 0; i < n; ++i)
    {
        const float *p = pts + static_cast<std::size_t>(i) * stride;
        const float one = 1.0F, intensity = stride > 3 ? p[3] : 0.0F;
        std::uint8_t *rec = msg.data.data() + static_cast<std::size_t>(i) * 32;
        std::memcpy(rec + 0, p, 12);
        std::memcpy(rec + 12, &one, 4);
        std::memcpy(rec + 16, &intensity, 4);
    }
    return msg;
}

void on_spin()
{
    using rclcpp::shim::set_sink;
    set_sink<PointCloud2>("ground_pointcloud", [](const PointCloud2 &m) {
        g_out.ground = m.data;
        g_out.ground_step = m.point_step;
        g_out.has_ground = true;
    });
    set_sink<PointCloud2>("obstacle_pointcloud", [](const PointCloud2 &m) {
        g_out.obstacle = m.data;
        g_out.obstacle_step = m.point_step;
        g_out.has_obstacle = true;
    });
    set_sink<PointCloud2>("clustered_pointcloud", [](const PointCloud2 &m) {
        g_out.clustered = m.data;
        g_out.clustered_step = m.point_step;
        g_out.has_clustered = true;
    });
    set_sink<MarkerArray>("polygonization", [](const MarkerArray &m) {
        g_out.has_markers = true;
        for (const auto &mk : m.markers)
        {
            g_out.marker_sizes.push_back(static_cast<std::uint32_t>(mk.points.size()));
            g_out.marker_ids.push_back(mk.id);
            for (const auto &p : mk.points)
            {
                g_out.marker_xyz.push_back(p.x);
                g_out.marker_xyz.push_back(p.y);
                g_out.marker_xyz.push_back(p.z);
            }
        }
    });
    rclcpp::shim::deliver<PointCloud2>("pointcloud", make_message(g_points, g_n, g_stride));
}
} // namespace

extern "C"
{

// Runs one frame through a fresh reference node. rand_seed seeds std::rand() before the node draws the cluster colours
// (conversions.cpp:48-50). Returns 0, or 1 when the node's main reported an exception (text via ref_node_error()).
int ref_node_run(const float *pts, std::uint32_t n, std::uint32_t stride_floats, unsigned rand_seed)
{
    g_out = Captured{};
    g_points = pts;
    g_n = n;
    g_stride = stride_floats;
    rclcpp::shim::spin_hook() = on_spin;
    std::srand(rand_seed);
    int rc = 1;
    try
    {
        rc = ref_processor_main(0, nullptr);
    }
    catch (const std::exception &e)
    {
        g_error = e.what();
    }
    return rc == EXIT_SUCCESS ? 0 : 1;
}

// sizes: [0] ground bytes, [1] obstacle bytes, [2] clustered bytes, [3] markers, [4] marker points,
//        [5..7] point_step of the three clouds, [8] bit mask of the topics that were published
void ref_node_sizes(std::uint64_t *sizes)
{
    sizes[0] = g_out.ground.size();
    sizes[1] = g_out.obstacle.size();
    sizes[2] = g_out.clustered.size();
    sizes[3] = g_out.marker_sizes.size();
    sizes[4] = g_out.marker_xyz.size() / 3;
    sizes[5] = g_out.ground_step;
    sizes[6] = g_out.obstacle_step;
    sizes[7] = g_out.clustered_step;
    sizes[8] = (g_out.has_ground ? 1U : 0U) | (g_out.has_obstacle ? 2U : 0U) | (g_out.has_clustered ? 4U : 0U) |
               (g_out.has_markers ? 8U : 0U);
}

void ref_node_fetch(std::uint8_t *ground, std::uint8_t *obstacle, std::uint8_t *clustered, std::uint32_t *marker_sizes,
                    std::int32_t *marker_ids, double *marker_xyz)
{
    if (ground && !g_out.ground.empty())
        std::memcpy(ground, g_out.ground.data(), g_out.ground.size());
    if (obstacle && !g_out.obstacle.empty())
        std::memcpy(obstacle, g_out.obstacle.data(), g_out.obstacle.size());
    if (clustered && !g_out.clustered.empty())
        std::memcpy(clustered, g_out.clustered.data(), g_out.clustered.size());
    if (marker_sizes && !g_out.marker_sizes.empty())
        std::memcpy(marker_sizes, g_out.marker_sizes.data(), g_out.marker_sizes.size() * 4);
    if (marker_ids && !g_out.marker_ids.empty())
        std::memcpy(marker_ids, g_out.marker_ids.data(), g_out.marker_ids.size() * 4);
    if (marker_xyz && !g_out.marker_xyz.empty())
        std::memcpy(marker_xyz, g_out.marker_xyz.data(), g_out.marker_xyz.size() * 8);
}

} // extern "C"
