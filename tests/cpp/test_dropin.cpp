// C++ parity harness for the drop-in classes: does what Processor::process does around the hot path
// (reference src/processor.cpp:150-195) with the replacement Segmenter / Clusterer headers, then
// dumps the results for tests/test_gpu_dropin.py to compare with the oracle.
//   usage: test_dropin <points.f32 (N x 4 floats)> <out_prefix>
#include "clustering.hpp"
#include "segmentation.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <vector>

using namespace lidar_processing;

template <typename T> void dump(const std::string &path, const std::vector<T> &v)
{
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
}

int main(int argc, char **argv)
{
    if (argc != 3)
        return 2;
    std::ifstream in(argv[1], std::ios::binary | std::ios::ate);
    const std::size_t bytes = static_cast<std::size_t>(in.tellg());
    in.seekg(0);
    std::vector<float> raw(bytes / 4);
    in.read(reinterpret_cast<char *>(raw.data()), static_cast<std::streamsize>(bytes));
    const std::size_t n = raw.size() / 4;

    pcl::PointCloud<pcl::PointXYZI> cloud_in;
    for (std::size_t i = 0; i < n; ++i)
        cloud_in.emplace_back(raw[i * 4 + 0], raw[i * 4 + 1], raw[i * 4 + 2], raw[i * 4 + 3]);

    try
    {
        Segmenter segmenter;
        Clusterer clusterer;
        std::vector<SegmentationLabel> segmentation_labels;
        pcl::PointCloud<pcl::PointXYZI> ground_points, obstacle_points;
        // twice, like consecutive callbacks reusing the member vectors (processor.cpp:126-132)
        for (int rep = 0; rep < 2; ++rep)
            segmenter.segment(cloud_in, segmentation_labels, ground_points, obstacle_points);

        pcl::PointCloud<pcl::PointXYZRGBL> obstacle_cloud; // processor.cpp:158-163
        obstacle_cloud.reserve(obstacle_points.size());
        for (const auto &p : obstacle_points)
            obstacle_cloud.emplace_back(p.x, p.y, p.z, 0, 255, 0, 1);

        std::vector<ClusteringLabel> cluster_labels;
        clusterer.cluster(obstacle_cloud, cluster_labels);
        for (const auto label : cluster_labels)
            if (label == Clusterer::UNDEFINED)
                throw std::runtime_error("Undefined label found (clustering)"); // processor.cpp:186-189

        // the caller's split by label (processor.cpp:180-200), once on the host as the reference does it and
        // once through the device-side extension; both must agree point for point
        std::vector<pcl::PointCloud<pcl::PointXYZ>> host_split, device_split;
        if (!cluster_labels.empty())
        {
            const auto max_label = *std::max_element(cluster_labels.cbegin(), cluster_labels.cend());
            host_split.resize(max_label + 1);
            for (std::size_t i = 0; i < obstacle_cloud.size(); ++i)
                if (cluster_labels[i] != Clusterer::INVALID)
                    host_split[cluster_labels[i]].emplace_back(obstacle_cloud.points[i].x, obstacle_cloud.points[i].y,
                                                               obstacle_cloud.points[i].z);
            host_split.erase(std::remove_if(host_split.begin(), host_split.end(),
                                            [](const pcl::PointCloud<pcl::PointXYZ> &c) { return c.empty(); }),
                             host_split.end());
        }
        clusterer.split_last_clusters(device_split);
        if (device_split.size() != host_split.size())
            return 4;
        for (std::size_t k = 0; k < host_split.size(); ++k)
        {
            if (device_split[k].size() != host_split[k].size())
                return 5;
            if (std::memcmp(device_split[k].points.data(), host_split[k].points.data(),
                            host_split[k].size() * sizeof(pcl::PointXYZ)) != 0)
                return 6;
        }

        // outlines of those clusters on the device, dumped for the comparison with the reference's host functions
        for (int policy = 0; policy < 3; ++policy)
        {
            std::vector<std::vector<Clusterer::OutlinePoint>> outlines;
            std::vector<std::uint32_t> host_ids;
            clusterer.outline_last_clusters(static_cast<Clusterer::OutlinePolicy>(policy), outlines, host_ids);
            if (outlines.size() != device_split.size())
                return 7;
            std::vector<std::uint32_t> sizes;
            std::vector<float> xy;
            for (const auto &o : outlines)
            {
                sizes.push_back(static_cast<std::uint32_t>(o.size()));
                for (const auto &v : o)
                    xy.insert(xy.end(), {v.x, v.y});
            }
            const std::string tag = std::string(argv[2]) + (policy == 0 ? ".convex" : (policy == 1 ? ".concave_small" : ".concave"));
            dump(tag + ".sizes.u32", sizes);
            dump(tag + ".xy.f32", xy);
            dump(tag + ".host_ids.u32", host_ids);
        }

        // colourised cloud: same std::rand() sequence on both sides (conversions.cpp:32-60)
        {
            std::srand(1234U);
            pcl::PointCloud<pcl::PointXYZRGB> want;
            for (const auto &cluster : host_split)
            {
                const auto r = static_cast<std::uint8_t>(std::rand() % 256);
                const auto g = static_cast<std::uint8_t>(std::rand() % 256);
                const auto b = static_cast<std::uint8_t>(std::rand() % 256);
                for (const auto &point : cluster.points)
                    want.push_back(pcl::PointXYZRGB(point.x, point.y, point.z, r, g, b));
            }
            std::srand(1234U);
            pcl::PointCloud<pcl::PointXYZRGB> got;
            clusterer.colorize_last_clusters(got);
            if (got.size() != want.size())
                return 8;
            for (std::size_t i = 0; i < want.size(); ++i)
                if (std::memcmp(&got.points[i], &want.points[i], 20) != 0) // x, y, z, 1.0f, b g r a
                    return 9;
        }
        // marker strips of the convex outlines (conversions.hpp:72-120): closed, z = 0, empty outlines skipped
        {
            std::vector<std::vector<Clusterer::OutlinePoint>> outlines;
            std::vector<std::uint32_t> host_ids, marker_ids;
            clusterer.outline_last_clusters(Clusterer::OutlinePolicy::CONVEX, outlines, host_ids);
            std::vector<std::vector<Clusterer::MarkerPoint>> strips;
            clusterer.marker_points_of_last_outlines(strips, marker_ids);
            std::size_t at = 0;
            for (std::size_t k = 0; k < outlines.size(); ++k)
            {
                if (outlines[k].empty())
                    continue;
                if (at >= strips.size() || marker_ids[at] != k || strips[at].size() != outlines[k].size() + 1U)
                    return 10;
                for (std::size_t v = 0; v <= outlines[k].size(); ++v)
                {
                    const auto &o = outlines[k][v == outlines[k].size() ? 0 : v];
                    const auto &m = strips[at][v];
                    if (m.x != static_cast<double>(o.x) || m.y != static_cast<double>(o.y) || m.z != 0.0)
                        return 11;
                }
                ++at;
            }
            if (at != strips.size())
                return 12;
        }

        std::vector<std::uint32_t> seg(segmentation_labels.size());
        for (std::size_t i = 0; i < seg.size(); ++i)
            seg[i] = static_cast<std::uint32_t>(segmentation_labels[i]);
        std::vector<float> ground_xyz, obstacle_xyz;
        for (const auto &p : ground_points)
            ground_xyz.insert(ground_xyz.end(), {p.x, p.y, p.z, p.intensity});
        for (const auto &p : obstacle_points)
            obstacle_xyz.insert(obstacle_xyz.end(), {p.x, p.y, p.z, p.intensity});
        const std::string prefix = argv[2];
        dump(prefix + ".seg.u32", seg);
        dump(prefix + ".ground.f32", ground_xyz);
        dump(prefix + ".obstacle.f32", obstacle_xyz);
        dump(prefix + ".clusters.i32", cluster_labels);

        // configuration errors surface as exceptions, empty clouds are fine
        bool threw = false;
        try
        {
            SegmentationConfiguration bad;
            bad.number_of_planar_partitions = 0U;
            segmenter.update_configuration(bad);
        }
        catch (const std::invalid_argument &)
        {
            threw = true;
        }
        pcl::PointCloud<pcl::PointXYZ> empty_in, g, o;
        std::vector<SegmentationLabel> l;
        segmenter.segment(empty_in, l, g, o);
        std::vector<ClusteringLabel> cl;
        clusterer.cluster(empty_in, cl);
        if (!threw || !l.empty() || !cl.empty())
            return 3;
        std::printf("ok %zu points, %zu ground, %zu obstacle, %d clusters, %zu split clouds\n", n, ground_points.size(),
                    obstacle_points.size(),
                    cluster_labels.empty() ? 0 : 1 + *std::max_element(cluster_labels.begin(), cluster_labels.end()),
                    device_split.size());
    }
    catch (const std::exception &e)
    {
        std::cerr << "exception: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
