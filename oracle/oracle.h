/* TEST INFRASTRUCTURE ONLY — CPU oracle for the ground-segmentation + Fast-Euclidean-Clustering
 * hot path of YevgeniyEngineer/LiDAR-Processing. Never linked into, imported by or called from the
 * product path (lidar-processing_b200/). Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may use it.
 *
 * Parity pinning status
 *  - oracle_cluster / oracle_kd_order: PINNED against the unmodified reference Clusterer / KDTree
 *    compiled from /root/reference/src (oracle/_ref/libref_cluster.so) on all 154 data frames and on
 *    synthetic tie-heavy clouds (tests/test_oracle_pinning.py; generated goldens in tests/golden/).
 *  - oracle_segment: "parity unpinned" at the Eigen boundary. The reference's segmentation.cpp needs
 *    Eigen 3.4 + PCL, neither vendored nor installed, and the reference ships no test or golden
 *    vector for it. The restatement follows segmentation.cpp line by line and restates Eigen 3.4's
 *    JacobiSVD (square real case) from its published algorithm; it is checked against numpy's
 *    float64 eigensolver (normal direction, sign convention) only.
 */
#ifndef LIDAR_B200_ORACLE_H
#define LIDAR_B200_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    /* reference: src/segmentation.hpp:48-56 */
    typedef struct oracle_seg_cfg
    {
        float sensor_height_m;
        float orthogonal_distance_threshold;
        float initial_seed_threshold;
        uint32_t number_of_iterations;
        uint32_t number_of_planar_partitions;
        uint32_t number_of_lower_point_representatives;
    } oracle_seg_cfg;

    /* reference: src/clustering.hpp:42-48 */
    typedef struct oracle_clu_cfg
    {
        float distance_squared;
        float cluster_quality;
        uint32_t min_cluster_size;
        uint32_t max_cluster_size;
    } oracle_clu_cfg;

    void oracle_seg_cfg_default(oracle_seg_cfg *cfg);
    void oracle_clu_cfg_default(oracle_clu_cfg *cfg);

    /* Segmenter::segment (src/segmentation.cpp:311-345).
     * pts: n records of stride_floats floats, xyz first. tie_mode 0: std::sort on x (what the
     * reference compiles to here: serial PSTL backend); 1: stable order (x, then original index).
     * labels: n entries, IN/OUT — only classified points are written (the reference's
     * labels.resize() keeps old entries, segmentation.cpp:315). ground_idx / obstacle_idx: original
     * indices in the reference's output-cloud order. planes_out (optional): P * iterations * 4 floats
     * (a,b,c,d), NaN where no fit ran. seg_status_out (optional, P entries): 0 ok, 1 "<3 points"
     * (left UNKNOWN), 2 "Failed ground segmentation" (all OBSTACLE). Returns 0, or -1 on bad config. */
    int oracle_segment(const float *pts, uint32_t n, uint32_t stride_floats, const oracle_seg_cfg *cfg, int tie_mode,
                       uint32_t *labels, uint32_t *ground_idx, uint32_t *n_ground, uint32_t *obstacle_idx,
                       uint32_t *n_obstacle, float *planes_out, int32_t *seg_status_out);

    /* Restated Eigen 3.4 JacobiSVD<Matrix3f>(ComputeThinV) on a row-major 3x3; v_out row-major 3x3
     * (columns sorted by descending singular value), sv_out 3 singular values. Returns sweeps, -1 if
     * the input is not finite. */
    int oracle_jacobi_svd3(const float *a, float *v_out, float *sv_out);

    /* Clusterer::cluster (src/clustering.cpp:47-125) restated with its own k-d tree
     * (kdtree.hpp:174-225, 292-341) built with the real std::nth_element. labels_out: m int32. */
    int oracle_cluster(const float *pts, uint32_t m, uint32_t stride_floats, const oracle_clu_cfg *cfg,
                       int32_t *labels_out);

    /* k-d tree pre-order: order_out[rank] = point index. mode 0: std::nth_element (ground truth of
     * this toolchain's libstdc++); 1: hand transcription of libstdc++ 13 introselect
     * (bits/stl_algo.h:1871-1978, stl_heap.h) — what the device emulates. */
    int oracle_kd_order(const float *pts, uint32_t m, uint32_t stride_floats, int mode, uint32_t *order_out);

    /* Same, also returning the permuted node array (m entries: point index at each array slot). */
    int oracle_kd_build(const float *pts, uint32_t m, uint32_t stride_floats, int mode, uint32_t *slot_to_index_out);

    /* Connected components of the graph {d2(i,j) <= distance_squared} with d2 computed exactly as
     * KDTree::dist_sqr (kdtree.hpp:145-163). root_out[i] = minimum point index of i's component. */
    int oracle_cc(const float *pts, uint32_t m, uint32_t stride_floats, float distance_squared, uint32_t *root_out);

    /* Model of the device formulation (test aid): voxel-grid neighbour sets, hits ordered by the
     * given pre-order rank, de-duplicated FIFO, one independent replay per r-component, then a
     * label-compaction pass. Must equal oracle_cluster bit for bit. */
    int oracle_cluster_model(const float *pts, uint32_t m, uint32_t stride_floats, const oracle_clu_cfg *cfg,
                             const uint32_t *rank_of_point, int32_t *labels_out, uint64_t *stats_out /* 8 */);

    /* Canonicalise a label vector: clusters renumbered by ascending minimum member index, negative
     * labels preserved. Returns the number of clusters. */
    uint32_t oracle_canonicalise(const int32_t *labels, uint32_t m, int32_t *out);

    /* FNV-1a 64 over a byte range (golden fingerprints). */
    uint64_t oracle_fnv1a64(const void *data, uint64_t nbytes);

#ifdef __cplusplus
}
#endif

#endif /* LIDAR_B200_ORACLE_H */
