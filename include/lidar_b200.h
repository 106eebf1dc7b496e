/* lidar_b200 — C ABI of the B200-native ground-segmentation + Fast-Euclidean-Clustering path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. The C++ classes
 * lidar_processing::Segmenter / lidar_processing::Clusterer (lidar-processing_b200/dropin/
 * segmentation.hpp, clustering.hpp — same public interface as the reference headers) are thin shells
 * over these entry points; INTEGRATION.md shows the binding a maintainer of the reference adds.
 *
 * The library owns pinned host staging buffers, device arenas and CUDA streams. All functions return
 * 0 on success and a non-zero lidar_b200_status otherwise; lidar_b200_last_error() gives the text.
 * A context is not thread-safe (the reference classes are used from one executor thread,
 * reference src/processor.cpp:279); use one context per thread / per GPU.
 */
#ifndef LIDAR_B200_H
#define LIDAR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

    typedef struct lidar_b200_ctx lidar_b200_ctx;

    typedef enum lidar_b200_status
    {
        LIDAR_B200_OK = 0,
        LIDAR_B200_ERR_CUDA = 1,        /* a CUDA call failed (no analogue in the reference; the C++ shells throw) */
        LIDAR_B200_ERR_INVALID = 2,     /* bad argument */
        LIDAR_B200_ERR_UNSUPPORTED = 3, /* configuration outside the supported envelope */
        LIDAR_B200_ERR_CAPACITY = 4,    /* batch larger than the reserved arenas and growth failed */
        LIDAR_B200_ERR_INPUT = 5        /* non-finite / out-of-range coordinates (unspecified in the reference); outlines: a cluster on which the reference does not terminate */
    } lidar_b200_status;

    /* replaces lidar_processing::SegmentationConfiguration (reference src/segmentation.hpp:48-56) */
    typedef struct lidar_b200_seg_cfg
    {
        float sensor_height_m;                          /* 1.73 */
        float orthogonal_distance_threshold;            /* 0.3  */
        float initial_seed_threshold;                   /* 0.6  */
        uint32_t number_of_iterations;                  /* 3    */
        uint32_t number_of_planar_partitions;           /* 2    */
        uint32_t number_of_lower_point_representatives; /* 5000 */
    } lidar_b200_seg_cfg;

    /* replaces lidar_processing::ClusteringConfiguration (reference src/clustering.hpp:42-48) */
    typedef struct lidar_b200_clu_cfg
    {
        float distance_squared;    /* 0.18 */
        float cluster_quality;     /* 0.5  */
        uint32_t min_cluster_size; /* 4    */
        uint32_t max_cluster_size; /* UINT32_MAX */
    } lidar_b200_clu_cfg;

    /* SegmentationLabel values (reference src/segmentation.hpp:41-46) */
#define LIDAR_B200_SEG_UNKNOWN 0u
#define LIDAR_B200_SEG_GROUND 1u
#define LIDAR_B200_SEG_OBSTACLE 2u
    /* ClusteringLabel specials (reference src/clustering.hpp:53-54) */
#define LIDAR_B200_CLU_UNDEFINED INT32_MIN
#define LIDAR_B200_CLU_INVALID (-1)

    void lidar_b200_seg_cfg_default(lidar_b200_seg_cfg *cfg);
    void lidar_b200_clu_cfg_default(lidar_b200_clu_cfg *cfg);

    /* replaces Segmenter::Segmenter() / Clusterer::Clusterer() + reserve_memory(200'000)
     * (reference src/segmentation.cpp:33-60, src/clustering.cpp:27-45). `max_points` is the total
     * number of points of one batch the arenas are sized for up front (they grow on demand),
     * `max_frames` the number of frames per batch. Fails when the CUDA library/device is missing —
     * there is no CPU fallback. */
    int lidar_b200_create(int device, uint32_t max_points, uint32_t max_frames, lidar_b200_ctx **ctx_out);
    void lidar_b200_destroy(lidar_b200_ctx *ctx);

    /* replaces Segmenter::reserve_memory / Clusterer::reserve_memory (segmentation.hpp:66, clustering.hpp:61) */
    int lidar_b200_reserve(lidar_b200_ctx *ctx, uint32_t max_points, uint32_t max_frames);

    /* replaces Segmenter::update_configuration / Clusterer::update_configuration */
    int lidar_b200_seg_configure(lidar_b200_ctx *ctx, const lidar_b200_seg_cfg *cfg);
    int lidar_b200_clu_configure(lidar_b200_ctx *ctx, const lidar_b200_clu_cfg *cfg);

    /* replaces Segmenter::segment<PointT> (reference src/segmentation.cpp:311-345).
     * points: n records, `stride_bytes` apart (16 for PointXYZ, 32 for PointXYZI), x,y,z = the first
     * three floats of each record. labels_inout: n entries; classified points receive GROUND/OBSTACLE,
     * all others keep their previous value (the reference's labels.resize() does not reset old
     * entries, segmentation.cpp:315). ground_idx_out / obstacle_idx_out: capacity n each; original
     * point indices in the order the reference push_back()s them into ground_cloud / obstacle_cloud. */
    int lidar_b200_segment(lidar_b200_ctx *ctx, const void *points, uint32_t n, uint32_t stride_bytes,
                           uint32_t *labels_inout, uint32_t *ground_idx_out, uint32_t *n_ground_out,
                           uint32_t *obstacle_idx_out, uint32_t *n_obstacle_out);

    /* replaces Clusterer::cluster<PointT> (reference src/clustering.cpp:47-125). labels_out: m int32:
     * 0..K-1 dense cluster ids, -1 INVALID. */
    int lidar_b200_cluster(lidar_b200_ctx *ctx, const void *points, uint32_t m, uint32_t stride_bytes,
                           int32_t *labels_out);

    /* ---- batched frame pipeline: segment + cluster of many independent frames per call ----------
     * (the throughput path; one call replaces a sequence of Processor::process iterations,
     * reference src/processor.cpp:150-178, with the obstacle cloud staying on the device).
     *
     * stage : host -> pinned staging -> device (asynchronous on the context's stream)
     * run   : all kernels for the staged batch (asynchronous)
     * fetch : device -> host results; blocks until they are there.
     * Frame f of a fetched batch owns [point_offset[f], point_offset[f] + n_points[f]) in the
     * concatenated output arrays; its obstacle-indexed outputs (obstacle_idx, cluster_labels) use the
     * same offset with n_obstacle[f] live entries. */
    int lidar_b200_batch_stage(lidar_b200_ctx *ctx, uint32_t n_frames, const void *const *points,
                               const uint32_t *n_points, uint32_t stride_bytes);
    int lidar_b200_batch_run(lidar_b200_ctx *ctx);
    int lidar_b200_batch_fetch(lidar_b200_ctx *ctx, uint32_t *point_offset_out /* [n_frames] */,
                               uint32_t *seg_labels_out /* [sum n] */, uint32_t *ground_idx_out /* [sum n] */,
                               uint32_t *n_ground_out /* [n_frames] */, uint32_t *obstacle_idx_out /* [sum n] */,
                               uint32_t *n_obstacle_out /* [n_frames] */, int32_t *cluster_labels_out /* [sum n] */,
                               uint32_t *n_clusters_out /* [n_frames] */);
    /* fetch split in two: _fetch_async enqueues the device -> host copies behind the kernels and returns;
     * _wait blocks until the results are in the caller's arrays (lidar_b200_batch_fetch = both). The
     * arrays must stay valid until _wait (or the next _stage on this context, which waits first). */
    int lidar_b200_batch_fetch_async(lidar_b200_ctx *ctx, uint32_t *point_offset_out, uint32_t *seg_labels_out,
                                     uint32_t *ground_idx_out, uint32_t *n_ground_out, uint32_t *obstacle_idx_out,
                                     uint32_t *n_obstacle_out, int32_t *cluster_labels_out, uint32_t *n_clusters_out);
    int lidar_b200_batch_wait(lidar_b200_ctx *ctx);
    /* waits for the stream; used by benchmarks that keep inputs and outputs resident in HBM */
    int lidar_b200_sync(lidar_b200_ctx *ctx);

    /* ---- per-cluster point compaction on the device (the step right after Clusterer::cluster in the
     * reference's caller): replaces the host loop of Processor::process that splits the obstacle cloud
     * by label, skips INVALID points and erases empty clouds (reference src/processor.cpp:180-200).
     *
     * _group_clusters: enqueue behind the last lidar_b200_batch_run / lidar_b200_cluster of this context.
     * _fetch_clusters: blocks; per frame f (O = point_offset[f], K = n_clusters[f]):
     *   cluster_offset_out[O + f + k], k = 0..K : CSR offsets, cluster k owns grouped points
     *       [O + offset[k], O + offset[k+1]); offset[K] = number of points in a valid cluster
     *   cluster_points_out : 4 floats per point = pcl::PointXYZ(x, y, z) records (data[3] = 1.0f), in
     *       ascending obstacle-cloud index inside a cluster = the reference's emplace_back order
     *   cluster_point_idx_out : the obstacle-cloud index of every grouped point (optional, may be NULL)
     * Array sizes: n_clusters_out [n_frames], cluster_offset_out [sum n + n_frames],
     * cluster_points_out [sum n][4], cluster_point_idx_out [sum n] (sum n over the padded frame slots,
     * i.e. point_offset of the last frame + its n). */
    int lidar_b200_batch_group_clusters(lidar_b200_ctx *ctx);
    int lidar_b200_batch_fetch_clusters(lidar_b200_ctx *ctx, uint32_t *n_clusters_out, uint32_t *cluster_offset_out,
                                        float *cluster_points_out, uint32_t *cluster_point_idx_out);

    /* ---- ordered convex outlines per cluster on the device (the polygonization step that follows the
     * split in the reference's caller, src/processor.cpp:210-214). Runs on the grouped clusters of
     * lidar_b200_batch_group_clusters and re-enacts, bit for bit in float32:
     *   LIDAR_B200_HULL_CONVEX        findOrderedConvexOutlines (reference src/polygon_simplification.cpp:31-79):
     *       geom::constructConvexHull with ANDREW_MONOTONE_CHAIN up to 1000 points and CHAN above
     *       (reference Convex-Hull/convex_hull.hpp:212-281, 366-424), COUNTERCLOCKWISE, open;
     *   LIDAR_B200_HULL_CONCAVE_SMALL the convex branch of findOrderedConcaveOutlines
     *       (src/polygon_simplification.cpp:100-118): clusters below 20 points. Clusters from 20 points
     *       on get 0 vertices here: their Delaunay-based concave hull stays on the host.
     *   LIDAR_B200_HULL_CONCAVE       findOrderedConcaveOutlines as a whole (src/polygon_simplification.cpp:81-149):
     *       that branch below 20 points, and from 20 points on geometry::ConcaveHull with chi = 0.2
     *       (reference Concave-Hull/concave_hull.hpp:96-193 over Concave-Hull/delaunator.cpp) re-enacted value for
     *       value in float64 — seed triangle, sweep order (std::sort by distance from its circumcentre), advancing
     *       hull, edge flips, erosion of the boundary edges by a max-heap of their lengths. These outlines are
     *       CLOSED like the reference's (the first vertex is repeated at the end).
     * _fetch_hulls: blocks; per frame f (O = point_offset[f], K = n_clusters[f]):
     *   hull_offset_out[O + f + k], k = 0..K : CSR offsets into the frame's vertex list; a cluster with an
     *       empty hull (fewer than 3 points) is the one the reference drops from its output vector
     *   hull_xy_out : 2 floats per vertex = geom::Point<float>{x, y}, frame f's vertices start at 2 * O
     *   hull_point_idx_out : obstacle-cloud index of every vertex (optional, may be NULL)
     *   n_vertices_out[f] = hull_offset[K]
     * Array sizes: n_vertices_out [n_frames], hull_offset_out [sum n + n_frames], hull_xy_out [sum n][2],
     * hull_point_idx_out [sum n]. Errors: LIDAR_B200_ERR_UNSUPPORTED for a cluster above ~1.04 M points,
     * LIDAR_B200_ERR_INPUT when the Jarvis march of a CHAN cluster does not close (duplicate hull vertices in
     * different subsets; the reference loops forever on such a cluster) and, in mode CONCAVE, for a cluster of 20 or
     * more points that are all collinear or all coincident in (x, y) (the reference throws "not triangulation",
     * delaunator.cpp:299, respectively reads out of bounds): the outputs are still complete, that cluster has 0
     * vertices and every other outline is valid; also when two different x or y values of a convex chain are less than
     * FLT_EPSILON apart (possible only for unquantised coordinates below 1 m): the reference's comparators are not a
     * strict weak order there, its own outline is unspecified, the one delivered uses exact comparisons.
     * LIDAR_B200_ERR_CAPACITY when the closed outlines of a frame need more
     * than one vertex per staged point of the frame (only possible when nearly every point is an outline vertex): the
     * outlines that do not fit have 0 vertices. hull_point_idx_out in mode CONCAVE: a point of the cluster with exactly
     * the vertex's (x, y); where several points coincide, the one with the lowest index unless the cluster's sweep
     * order had to be replayed (two different points exactly equally far from the seed circumcentre). */
#define LIDAR_B200_HULL_CONVEX 0u
#define LIDAR_B200_HULL_CONCAVE_SMALL 1u
#define LIDAR_B200_HULL_CONCAVE 2u
    int lidar_b200_batch_hull_outlines(lidar_b200_ctx *ctx, uint32_t mode);
    int lidar_b200_batch_fetch_hulls(lidar_b200_ctx *ctx, uint32_t *n_vertices_out, uint32_t *hull_offset_out,
                                     float *hull_xy_out, uint32_t *hull_point_idx_out);

    /* ---- output packing on the device: the two conversions the reference's caller runs before publishing
     * (reference src/processor.cpp:249-272).
     *
     * _fetch_colorized: convertClusteredCloudToColorizedCloud (reference src/conversions.cpp:32-60) of the grouped
     *   clusters, as the 32-byte pcl::PointXYZRGB records that convertPCLToPointCloud2 copies into the message
     *   (floats x, y, z, 1.0f; colour word b | g << 8 | r << 16 | 255 << 24 at byte 16; 12 padding bytes = 0).
     *   cluster_rgb: one word r << 16 | g << 8 | b per cluster, the frames' clusters end to end (n_rgb = sum of
     *   n_clusters) — the caller draws them the way the reference does (std::rand() % 256 three times per cluster,
     *   conversions.cpp:49-51), so the process-wide rand() sequence stays the caller's. colorized_out: 8 floats per
     *   record, frame f's records start at 8 * point_offset[f], cluster after cluster in push order
     *   (count = cluster_offset[K]). Size [sum n][8].
     * _fetch_marker_points: the points of convertPointXYZTypeToMarkerArray (reference src/conversions.hpp:72-120) for
     *   the outlines of lidar_b200_batch_hull_outlines: per non-empty outline its vertices as geometry_msgs::Point
     *   {double x, y, z = 0} plus the first vertex again (loop closure, :108-117). marker_points_out: 3 doubles per
     *   point, frame f's points start at 3 * 2 * point_offset[f]. marker_offset_out[O + f + k] = non-empty outlines
     *   before cluster k (k = 0..K): the marker of cluster k owns points [hull_offset[k] + marker_offset[k],
     *   hull_offset[k+1] + marker_offset[k+1]) and its marker.id is k when the outline vector keeps empty entries,
     *   marker_offset[k] after the reference's erase of empty outlines. n_markers_out[f] = markers of frame f.
     *   Sizes: marker_points_out [2 * sum n][3], marker_offset_out [sum n + n_frames], n_markers_out [n_frames]. */
    int lidar_b200_batch_fetch_colorized(lidar_b200_ctx *ctx, const uint32_t *cluster_rgb, uint64_t n_rgb,
                                         float *colorized_out);
    int lidar_b200_batch_fetch_marker_points(lidar_b200_ctx *ctx, uint32_t *n_markers_out, uint32_t *marker_offset_out,
                                             double *marker_points_out);

    /* Page-locked host memory for clouds and result arrays — the zero-copy counterpart of the
     * caller-owned cloud_in_ / label vectors of the reference (src/processor.cpp:123-126). Every entry
     * point accepts ordinary (pageable) host pointers and stages them through the library's own pinned
     * buffers; when a point array with 16-byte records or a result array lives in memory obtained here
     * (or from cudaMallocHost / cudaHostRegister) the copy engines read / write it directly and the
     * staging pass disappears. */
    int lidar_b200_host_alloc(void **ptr_out, uint64_t bytes);
    void lidar_b200_host_free(void *ptr);

    /* PCD v0.7 reader (host code): replaces pcl::io::loadPCDFile + Dataloader::convert on the feeding side
     * (reference src/dataloader.cpp:87-126, 139). Writes records `stride_bytes` apart: 16 = packed
     * (x, y, z, intensity); 32 = the pcl::PointXYZI / PointCloud2 wire layout the processor node receives
     * (x 0, y 4, z 8, 1.0f 12, intensity 16; conversions.cpp:62-85). Both are accepted as they are by
     * lidar_b200_segment / _batch_stage / _pipe_submit. DATA binary and DATA ascii; fields x y z required,
     * intensity optional (0), other fields skipped. points_out == NULL only reports *n_points_out.
     * error_out (optional) receives the reason on failure. Needs no CUDA device. */
    int lidar_b200_pcd_read(const char *path, void *points_out, uint64_t capacity_points, uint32_t stride_bytes,
                            uint64_t *n_points_out, char *error_out, uint32_t error_capacity);

    /* ---- frame pipeline: `depth` contexts on one GPU used round-robin, so that the upload of chunk
     * k+1, the kernels of chunk k and the download of chunk k-1 overlap. _submit = stage + run +
     * fetch_async of one chunk of frames on the next slot; the results of a chunk are complete once
     * `depth` further chunks have been submitted or after _drain. Same array conventions as the
     * batch calls; one host thread per pipe. */
    typedef struct lidar_b200_pipe lidar_b200_pipe;
    int lidar_b200_pipe_create(int device, uint32_t depth, uint32_t max_points, uint32_t max_frames,
                               lidar_b200_pipe **pipe_out);
    void lidar_b200_pipe_destroy(lidar_b200_pipe *pipe);
    int lidar_b200_pipe_seg_configure(lidar_b200_pipe *pipe, const lidar_b200_seg_cfg *cfg);
    int lidar_b200_pipe_clu_configure(lidar_b200_pipe *pipe, const lidar_b200_clu_cfg *cfg);
    int lidar_b200_pipe_submit(lidar_b200_pipe *pipe, uint32_t n_frames, const void *const *points,
                               const uint32_t *n_points, uint32_t stride_bytes, uint32_t *point_offset_out,
                               uint32_t *seg_labels_out, uint32_t *ground_idx_out, uint32_t *n_ground_out,
                               uint32_t *obstacle_idx_out, uint32_t *n_obstacle_out, int32_t *cluster_labels_out,
                               uint32_t *n_clusters_out);
    int lidar_b200_pipe_drain(lidar_b200_pipe *pipe);
    uint64_t lidar_b200_pipe_launch_count(const lidar_b200_pipe *pipe);
    /* Tells the pipe how many GPUs of this host run a pipe at the same time (the local world size). They share the host's
     * DMA path: from 3 GPUs on the pipe fetches results with one kernel that writes the used part of every result slot into
     * the caller's page-locked arrays (fewer bytes, mode 4) instead of slot-size copies by the copy engines (mode 0, the
     * default and the faster one at 1-2 GPUs). LIDAR_B200_FETCH_MODE overrides. */
    int lidar_b200_pipe_set_host_sharing(lidar_b200_pipe *pipe, uint32_t gpus_sharing_host);
    /* Result fetch mode the pipe is using (0 or 4, see above). */
    int lidar_b200_pipe_fetch_mode(const lidar_b200_pipe *pipe);
    const char *lidar_b200_pipe_last_error(const lidar_b200_pipe *pipe);

    /* diagnostics */
    /* plane coefficients (a,b,c,d) of every fit of the last batch: [frame][partition][iteration][4], NaN = no fit;
     * status per [frame][partition]: 0 ok, 1 "<3 points", 2 "Failed ground segmentation". */
    int lidar_b200_last_planes(lidar_b200_ctx *ctx, float *planes_out, int32_t *status_out);
    /* k-d pre-order rank of every point of the last cluster()/batch frame 0 (parity checks) */
    int lidar_b200_last_kd_rank(lidar_b200_ctx *ctx, uint32_t frame, uint32_t *rank_out, uint32_t capacity);
    /* component root (smallest member index) of every obstacle point of a frame of the last batch */
    int lidar_b200_last_cc_root(lidar_b200_ctx *ctx, uint32_t frame, uint32_t *root_out, uint32_t capacity);
    /* per-cluster timings of the concave outlines of the last lidar_b200_batch_hull_outlines(LIDAR_B200_HULL_CONCAVE)
     * (development aid; needs LIDAR_B200_CHI_STATS=1 in the environment when the context is created). 8 words per
     * task, largest clusters first, at most 4096 tasks: points, start ns, end ns (%globaltimer), cycles of the seed
     * search / the sort / the sweep / the erosion, triangles. */
    int lidar_b200_last_chi_stats(lidar_b200_ctx *ctx, uint64_t *stats_out, uint32_t capacity_tasks, uint32_t *n_tasks_out);
    /* per-component counters of the CTA-per-component replay of the last run (development aid; needs
     * LIDAR_B200_REPLAY_STATS=1 in the environment when the context is created). 8 words per job:
     * frame, members, kilo-cycles, rounds, direct rounds, entries taken, seeds, candidates scanned. */
    int lidar_b200_last_replay_stats(lidar_b200_ctx *ctx, uint32_t *stats_out, uint32_t capacity_jobs,
                                     uint32_t *n_jobs_out);
    /* number of kernels launched by this context so far */
    uint64_t lidar_b200_launch_count(const lidar_b200_ctx *ctx);
    /* single-frame lidar_b200_batch_run calls that went to the device as ONE CUDA-graph launch (the frame's kernels are
     * captured per launch geometry - the point count rounded up to 4096 - after a first launch-by-launch run; they are
     * still counted kernel by kernel in lidar_b200_launch_count). LIDAR_B200_GRAPH=0 in the environment turns it off. */
    uint64_t lidar_b200_graph_launch_count(const lidar_b200_ctx *ctx);
    /* elapsed GPU milliseconds between the start and the end of the last lidar_b200_batch_run (CUDA events) */
    int lidar_b200_last_run_ms(lidar_b200_ctx *ctx, float *ms_out);
    /* device time of a region of calls on this context (benchmarks): _region_begin records a CUDA event on the
     * context's stream, _region_end_ms records a second one behind everything enqueued since, waits for it and returns
     * the elapsed milliseconds. Lets a caller enqueue several lidar_b200_batch_run back to back (no host round trip
     * between them) and still time them on the device. */
    int lidar_b200_region_begin(lidar_b200_ctx *ctx);
    int lidar_b200_region_end_ms(lidar_b200_ctx *ctx, float *ms_out);
    /* per-stage GPU times of the last run, measured with CUDA events on the context's stream when
     * profiling is enabled. 9 stages: x-sort | gather+fit | compact | voxel grid | union-find |
     * component sort | k-d order | replay | label compaction. */
    int lidar_b200_set_profiling(lidar_b200_ctx *ctx, int enabled);
    int lidar_b200_last_stage_ms(lidar_b200_ctx *ctx, float *ms_out /* [9] */, uint32_t capacity);
    const char *lidar_b200_last_error(const lidar_b200_ctx *ctx);
    const char *lidar_b200_version(void);

#ifdef __cplusplus
}
#endif

#endif /* LIDAR_B200_H */
