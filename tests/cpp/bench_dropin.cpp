// Timing harness for the REAL drop-in call: the replacement Segmenter / Clusterer classes driven
// exactly the way Processor::process drives the reference's (reference src/processor.cpp:150-178) — one frame at a time,
// pcl::PointCloud<pcl::PointXYZI> in pageable memory (32-byte records), synchronous calls, the obstacle cloud rebuilt on
// the host as pcl::PointXYZRGBL between the two calls (processor.cpp:158-163).
//   usage: bench_dropin <frames.bin> [passes]
//   frames.bin: uint32 n_frames, uint32 counts[n_frames], then per frame counts[f] x 4 float32 (x y z intensity)
// Prints one JSON line: per-frame latency (segment + rebuild + cluster) p50 / p95 / mean over the last pass, frames/s.
#include "clustering.hpp"
#include "segmentation.hpp"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

using namespace lidar_processing;

int main(int argc, char **argv)
{
    if (argc < 2)
        return 2;
    const int passes = argc > 2 ? std::atoi(argv[2]) : 2;
    std::ifstream in(argv[1], std::ios::binary);
    std::uint32_t nf = 0;
    in.read(reinterpret_cast<char *>(&nf), 4);
    std::vector<std::uint32_t> counts(nf);
    in.read(reinterpret_cast<char *>(counts.data()), 4 * static_cast<std::streamsize>(nf));
    std::vector<pcl::PointCloud<pcl::PointXYZI>> clouds(nf);
    std::vector<float> raw;
    for (std::uint32_t f = 0; f < nf; ++f)
    {
        raw.resize(static_cast<std::size_t>(counts[f]) * 4);
        in.read(reinterpret_cast<char *>(raw.data()), static_cast<std::streamsize>(raw.size() * 4));
        clouds[f].reserve(counts[f]);
        for (std::uint32_t i = 0; i < counts[f]; ++i)
            clouds[f].emplace_back(raw[i * 4 + 0], raw[i * 4 + 1], raw[i * 4 + 2], raw[i * 4 + 3]);
    }
    if (!in)
    {
        std::cerr << "short input file" << std::endl;
        return 3;
    }
    try
    {
        Segmenter segmenter; // long-lived members of the node (processor.cpp:129-132)
        Clusterer clusterer;
        std::vector<SegmentationLabel> segmentation_labels;
        pcl::PointCloud<pcl::PointXYZI> ground_points, obstacle_points;
        pcl::PointCloud<pcl::PointXYZRGBL> obstacle_cloud;
        std::vector<ClusteringLabel> cluster_labels;
        std::vector<double> ms(nf), seg_ms(nf), clu_ms(nf);
        std::uint64_t clusters = 0, obstacles = 0;
        double wall_last = 0.0;
        for (int pass = 0; pass < passes; ++pass)
        {
            clusters = obstacles = 0;
            const auto w0 = std::chrono::steady_clock::now();
            for (std::uint32_t f = 0; f < nf; ++f)
            {
                const auto t0 = std::chrono::steady_clock::now();
                segmenter.segment(clouds[f], segmentation_labels, ground_points, obstacle_points);
                const auto t1 = std::chrono::steady_clock::now();
                obstacle_cloud.clear();
                obstacle_cloud.reserve(obstacle_points.size());
                for (const auto &p : obstacle_points)
                    obstacle_cloud.emplace_back(p.x, p.y, p.z, 0, 255, 0, 1);
                const auto t2 = std::chrono::steady_clock::now();
                clusterer.cluster(obstacle_cloud, cluster_labels);
                const auto t3 = std::chrono::steady_clock::now();
                ms[f] = std::chrono::duration<double, std::milli>(t3 - t0).count();
                seg_ms[f] = std::chrono::duration<double, std::milli>(t1 - t0).count();
                clu_ms[f] = std::chrono::duration<double, std::milli>(t3 - t2).count();
                obstacles += obstacle_cloud.size();
                std::int32_t mx = -1;
                for (const auto l : cluster_labels)
                    mx = std::max(mx, l);
                clusters += static_cast<std::uint64_t>(mx + 1);
            }
            wall_last = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
        }
        auto pct = [](std::vector<double> v, double q) {
            std::sort(v.begin(), v.end());
            return v[static_cast<std::size_t>(q * static_cast<double>(v.size() - 1))];
        };
        std::printf("{\"frames\": %u, \"passes\": %d, \"p50_ms\": %.4f, \"p95_ms\": %.4f, \"mean_ms\": %.4f, "
                    "\"segment_p50_ms\": %.4f, \"cluster_p50_ms\": %.4f, \"frames_per_s\": %.2f, "
                    "\"obstacle_points\": %llu, \"clusters\": %llu}\n",
                    nf, passes, pct(ms, 0.5), pct(ms, 0.95), wall_last * 1e3 / nf, pct(seg_ms, 0.5), pct(clu_ms, 0.5),
                    nf / wall_last, static_cast<unsigned long long>(obstacles), static_cast<unsigned long long>(clusters));
    }
    catch (const std::exception &e)
    {
        std::cerr << "exception: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
