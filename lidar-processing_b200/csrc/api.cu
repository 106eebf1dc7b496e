// C-ABI layer (include/lidar_b200.h): owns the CUDA stream, pinned staging buffers and device
// arenas, validates configurations and sequences the kernels of segment.cuh / cluster.cuh /
// kd_build.cuh / radix_sort.cuh for a batch of frames. No CPU fallback: every entry point either
// runs the CUDA path or returns an error.
#include "../../include/lidar_b200.h"

#include "cluster.cuh"
#include "common.cuh"
#include "group.cuh"
#include "hull.cuh"
#include "chi_shape.cuh"
#include "pack.cuh"
#include "kd_build.cuh"
#include "pcd_io.h"
#include "radix_sort.cuh"
#include "replay_cta.cuh"
#include "replay_cta2.cuh"
#include "replay_gen.cuh"
#include "segment.cuh"

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <chrono>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

using namespace lb;

namespace
{
constexpr uint32_t kGraphBucket = 4096u; // launch geometry of a single frame in flight, in points (see lidar_b200_batch_run)

template <typename T> struct DevBuf
{
    T *p{nullptr};
    size_t n{0};
};

template <typename T> struct PinBuf
{
    T *p{nullptr};
    size_t n{0};
};

uint32_t next_pow2(uint32_t v)
{
    uint32_t p = 1u;
    while (p < v)
        p <<= 1;
    return p;
}

uint32_t ceil_log2(uint32_t v)
{
    uint32_t b = 0u;
    while ((1ull << b) < v)
        ++b;
    return b;
}
} // namespace

struct lidar_b200_ctx
{
    int device{0};
    cudaStream_t stream{nullptr};
    cudaStream_t stream_big{nullptr}; // the CTA-per-component replay runs beside the warp-per-component one
    cudaStream_t stream_huge{nullptr}; // ... and the 1-CTA-per-SM launch for components beyond the normal state bitmap
    cudaStream_t stream_small{nullptr}; // ... and the short jobs of the window-synchronous replay beside the long ones
    cudaEvent_t ev_start{nullptr}, ev_stop{nullptr}, ev_fork{nullptr}, ev_join{nullptr}, ev_join2{nullptr}, ev_join3{nullptr}, ev_kd{nullptr};
    lidar_b200_seg_cfg seg_cfg{};
    lidar_b200_clu_cfg clu_cfg{};
    SegParams seg{};
    CluParams clu{};

    // reserved capacities
    uint32_t cap_pts{0}, cap_frames{0}, cap_table{0}, cap_planes{0};

    // pinned host staging
    PinBuf<float4> h_pts;
    PinBuf<uint32_t> h_u32[4]; // labels, ground idx, obstacle idx, cluster labels
    PinBuf<uint32_t> h_meta;   // off, cnt, toff, tcap, n_ground, n_obstacle, n_clusters : 7 * cap_frames
    PinBuf<uint32_t> h_err;    // device error flag of the batch being fetched

    // device arenas (per point)
    DevBuf<float4> d_pts, d_spts, d_obs, d_nodes, d_cpts, d_rpts;
    DevBuf<float4> d_ipts, d_mpts; // immutable records of the second-generation CTA replay (replay_cta2.cuh)
    DevBuf<uint32_t> d_rankpos;
    DevBuf<unsigned long long> d_mkey;
    int replay_version{5};         // LIDAR_B200_REPLAY_V: 1 = first-generation CTA replay (state in global memory),
                                   // 2 = second (state bitmap by member), 5 = window-synchronous CTA replay
                                   // (replay_gen.cuh), the default. (Generations 3 and 4 - live-candidate bitmaps by cell
                                   // order, a warp per component - measured slower than 2 and are gone, see profiles/README.md.)
    bool replay2_attr_done{false}, replay5_attr_done{false};
    bool frame_sort{true};         // LIDAR_B200_FRAME_SORT=0: the component sort uses the per-tile radix sort kernels for every batch size
    uint32_t replay5_big_threads{0};   // LIDAR_B200_REPLAY5_BIG_THREADS: 256 / 512 / 1024 threads per CTA for the long jobs;
                                       // 0 = by batch size: 512 up to 4 frames (the longest window chain is the latency of
                                       // a frame: p50 2.95 -> 2.53 ms), 256 beyond (a batch is bound by the total work and
                                       // wide CTAs only take registers from the other frames' jobs)
    uint32_t replay5_ctas_per_sm{3}; // LIDAR_B200_REPLAY5_CTAS_PER_SM
    DevBuf<uint32_t> d_nb27;       // packed neighbour table of the window-synchronous replay: first pos | count << 20, 27 per cell
    DevBuf<uint32_t> d_key_a, d_key_b, d_val_a, d_val_b, d_labels, d_gidx, d_oidx, d_slot_of, d_pos_of, d_parent,
        d_root, d_rank, d_gepos, d_lepos, d_state, d_seed_of, d_member_pos, d_queue, d_seed_label, d_comp_size, d_pslot;
    DevBuf<int32_t> d_clabels;
    DevBuf<unsigned long long> d_spill, d_pkey;
    DevBuf<uint8_t> d_flags, d_seed_valid;
    // per table slot
    DevBuf<unsigned long long> d_tkeys;
    DevBuf<uint32_t> d_tcount, d_tlive;
    DevBuf<uint4> d_cells;
    DevBuf<uint32_t> d_nbr;  // 27 neighbour cell ids per cell (rows live at cell-start positions)
    DevBuf<uint2> d_cinfo;   // {points, common parent} per cell
    DevBuf<uint2> d_biglist; // {frame, first member} of every component replayed by a whole CTA
    DevBuf<uint32_t> d_job_stats; // optional per-job counters of the CTA replay (LIDAR_B200_REPLAY_STATS=1)
    bool want_job_stats{false};
    // per frame
    DevBuf<uint32_t> d_meta; // off, cnt, toff, tcap, n_ground, n_obstacle, n_clusters, cursor : 8 * cap_frames
    DevBuf<uint32_t> d_err;
    DevBuf<float> d_planes;
    DevBuf<int32_t> d_status;
    DevBuf<uint32_t> d_hist;

    // current batch
    uint32_t n_frames{0}, total{0}, max_n{0}, max_tcap{0};
    bool batch_is_cluster_only{false};
    // what the last run_clustering worked on (input of lidar_b200_batch_group_clusters)
    const float4 *clu_pts{nullptr};
    const uint32_t *clu_counts{nullptr};
    uint32_t clu_max_m{0};
    bool grouped{false};
    DevBuf<uint32_t> d_goff; // CSR offsets of the grouped clusters: frame f at [off[f] + f, off[f] + f + K_f]
    // outlines of the grouped clusters (hull.cuh): CSR offsets in the layout of d_goff, vertices per frame, error bits
    DevBuf<uint32_t> d_hoff, d_hne, d_hnv, d_herr, d_rgb;
    DevBuf<unsigned char> d_chi;   // working sets of the concave outlines (chi_shape.cuh), 144 bytes per point slot
    DevBuf<uint32_t> d_chi_meta;   // [bucket counts 32 | bucket fill 32 | cursor]
    DevBuf<unsigned long long> d_chi_stats; // LIDAR_B200_CHI_STATS=1: 8 words per task for the first kChiStatTasks tasks
    bool chi_stats{false};
    uint32_t replay5_wide_lists{2}, replay5_wide_ctas_per_sm{2}; // LIDAR_B200_REPLAY5_WIDE_LISTS / _WIDE_CTAS_PER_SM
    uint32_t chi_ctas_per_sm{8}; // LIDAR_B200_CHI_CTAS_PER_SM: persistent CTAs of chi_outline_kernel per SM (4 are resident)
    uint32_t hull_mode{0};
    DevBuf<uint4> d_color;   // 32-byte PointXYZRGB records (pack.cuh), allocated on first use
    DevBuf<double> d_marker; // marker points, allocated on first use
    bool hulled{false}, hull_attr_done{false};
    std::vector<uint32_t> off, cnt;

    // optional per-stage CUDA-event timing (lidar_b200_set_profiling)
    static constexpr int kStages = 10;
    bool profiling{false};
    cudaEvent_t ev_stage[kStages + 1]{};
    int n_stage_marks{0};

    // outstanding asynchronous fetch (lidar_b200_batch_fetch_async .. lidar_b200_batch_wait)
    struct FetchReq
    {
        bool pending{false};
        uint32_t *point_offset{nullptr}, *n_ground{nullptr}, *n_obstacle{nullptr}, *n_clusters{nullptr};
        void *out[4]{nullptr, nullptr, nullptr, nullptr}; // seg labels, ground idx, obstacle idx, cluster labels
        bool direct[4]{false, false, false, false};       // destination is page-locked: DMA went straight into it
        bool payload_enqueued{false};                     // the result arrays are on their way (counts came first)
        bool with_worker{false}, worker_done{false};      // pipeline: the second phase belongs to the pipe's worker thread
        int worker_rc{0};
    } fetch;
    cudaEvent_t ev_region_a{nullptr}, ev_region_b{nullptr}; // lidar_b200_region_begin / _end_ms
    cudaEvent_t ev_counts{nullptr}; // per-frame counts of the batch are in h_meta
    std::vector<void *> copy_dst, copy_src; // copy list of the second fetch phase
    std::vector<size_t> copy_size;
    // One frame in flight: the ~58 launches of a frame (four streams, fork / join events) are captured into a CUDA graph
    // per launch geometry and replayed, so that the kernels of a 2 ms frame do not wait for the host to enqueue them.
    // The launch geometry of a single frame is rounded up to kGraphBucket points for that (kernels bound their loops by
    // the device-side counts). LIDAR_B200_GRAPH=0 turns it off.
    struct GraphEntry
    {
        uint32_t max_n{0};
        bool warm{false}, bad{false};
        cudaGraphExec_t exec{nullptr};
        uint64_t launches{0};
    };
    std::vector<GraphEntry> graphs;
    uint64_t epoch{0}, graphs_epoch{0}; // epoch: bumped by every (re)allocation and configuration change
    bool use_graph{true};
    uint32_t geom_total{0}; // c->total, or the single frame's bucketed size: sizes of the memsets / copies inside the run path
    uint64_t graph_launches{0};
    uint32_t stage_threads{4}; // LIDAR_B200_STAGE_THREADS: host threads that stage a batch of pageable clouds (1 = the caller's only)
    int fetch_mode{0}; // LIDAR_B200_FETCH_MODE (see profiles/README.md "result fetch modes"): 0 = one phase, full slots; 1 = two phases, full slots; 2 = exact sizes, plain copies; 3 = exact sizes, one batched call; 4 = one kernel writes the exact sizes straight into page-locked host memory (emit_results_kernel)

    uint32_t sm_count{148}, replay_ctas_per_sm{12}, replay_big_ctas_per_sm{3};
    uint64_t launches{0};
    float last_run_ms{0.0f};
    std::string err;

    uint32_t *m_off() { return d_meta.p; }
    uint32_t *m_cnt() { return d_meta.p + cap_frames; }
    uint32_t *m_toff() { return d_meta.p + 2 * static_cast<size_t>(cap_frames); }
    uint32_t *m_tcap() { return d_meta.p + 3 * static_cast<size_t>(cap_frames); }
    uint32_t *m_ng() { return d_meta.p + 4 * static_cast<size_t>(cap_frames); }
    uint32_t *m_no() { return d_meta.p + 5 * static_cast<size_t>(cap_frames); }
    uint32_t *m_nc() { return d_meta.p + 6 * static_cast<size_t>(cap_frames); }
    uint32_t *m_cursor() { return d_meta.p + 7 * static_cast<size_t>(cap_frames); }
};

namespace
{
int fail(lidar_b200_ctx *c, int code, const std::string &msg)
{
    if (c)
        c->err = msg;
    return code;
}

// device-visible address of page-locked host memory (mapped under unified addressing), or null
void *device_view_of_pinned(const void *p)
{
    if (!p)
        return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        (void)cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// true when `p` is page-locked host memory known to CUDA (lidar_b200_host_alloc, cudaMallocHost,
// cudaHostRegister): the copy engines can then read / write it directly, no staging pass needed
bool is_pinned(const void *p)
{
    if (!p)
        return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

#define LB_CUDA(c, call)                                                                                              \
    do                                                                                                                \
    {                                                                                                                 \
        const cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess)                                                                                        \
            return fail((c), LIDAR_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                \
    } while (0)

template <typename T> int dev_alloc(lidar_b200_ctx *c, DevBuf<T> &b, size_t n)
{
    if (b.n >= n && b.p)
        return 0;
    if (b.p)
        LB_CUDA(c, cudaFree(b.p));
    b.p = nullptr;
    b.n = 0;
    LB_CUDA(c, cudaMalloc(reinterpret_cast<void **>(&b.p), n * sizeof(T)));
    b.n = n;
    ++c->epoch; // device pointers changed: captured graphs are stale
    return 0;
}

template <typename T> int pin_alloc(lidar_b200_ctx *c, PinBuf<T> &b, size_t n)
{
    if (b.n >= n && b.p)
        return 0;
    if (b.p)
        LB_CUDA(c, cudaFreeHost(b.p));
    b.p = nullptr;
    b.n = 0;
    LB_CUDA(c, cudaMallocHost(reinterpret_cast<void **>(&b.p), n * sizeof(T)));
    b.n = n;
    return 0;
}

int reserve(lidar_b200_ctx *c, uint32_t pts, uint32_t frames)
{
    LB_CUDA(c, cudaSetDevice(c->device));
    pts = pts < 1024u ? 1024u : pts;
    frames = frames < 1u ? 1u : frames;
    const bool grow_pts = pts > c->cap_pts;
    const bool grow_frames = frames > c->cap_frames;
    if (grow_pts || grow_frames)
        LB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (grow_pts)
    {
        const size_t n = pts;
        int rc = 0;
        rc |= pin_alloc(c, c->h_pts, n);
        for (auto &h : c->h_u32)
            rc |= pin_alloc(c, h, n);
        rc |= dev_alloc(c, c->d_pts, n) | dev_alloc(c, c->d_spts, n) | dev_alloc(c, c->d_obs, n) |
              dev_alloc(c, c->d_nodes, n) | dev_alloc(c, c->d_cpts, n) | dev_alloc(c, c->d_rpts, n);
        DevBuf<uint32_t> *u32s[] = {&c->d_key_a,  &c->d_key_b,  &c->d_val_a,      &c->d_val_b, &c->d_labels,
                                    &c->d_gidx,   &c->d_oidx,   &c->d_slot_of,    &c->d_pos_of, &c->d_parent,
                                    &c->d_root,   &c->d_rank,   &c->d_gepos,      &c->d_lepos, &c->d_state,
                                    &c->d_seed_of, &c->d_member_pos, &c->d_queue, &c->d_seed_label, &c->d_comp_size, &c->d_pslot};
        for (auto *b : u32s)
            rc |= dev_alloc(c, *b, n);
        rc |= dev_alloc(c, c->d_clabels, n) | dev_alloc(c, c->d_spill, n) | dev_alloc(c, c->d_pkey, n) | dev_alloc(c, c->d_flags, n) |
              dev_alloc(c, c->d_seed_valid, n) | dev_alloc(c, c->d_nbr, 27u * n) | dev_alloc(c, c->d_cinfo, n) |
              dev_alloc(c, c->d_biglist, kBigListBuckets * (n / kCtaComponentMin + 1u)) | dev_alloc(c, c->d_ipts, n) |
              dev_alloc(c, c->d_mpts, n) | dev_alloc(c, c->d_rankpos, n) | dev_alloc(c, c->d_mkey, n) |
              dev_alloc(c, c->d_nb27, 27u * n);
        if (c->want_job_stats)
            rc |= dev_alloc(c, c->d_job_stats, 8u * (n / kCtaComponentMin + 1u));
        if (rc)
            return LIDAR_B200_ERR_CUDA;
        c->cap_pts = pts;
    }
    if (grow_frames)
    {
        if (dev_alloc(c, c->d_meta, 8 * static_cast<size_t>(frames) + 16) || pin_alloc(c, c->h_meta, 7 * static_cast<size_t>(frames)))
            return LIDAR_B200_ERR_CUDA;
        c->cap_frames = frames;
    }
    // hash-table slots: every frame reserves next_pow2(max(64, 2*n)) <= 4*n + 64 slots
    const size_t table = 4ull * c->cap_pts + 64ull * c->cap_frames;
    if (table > c->cap_table)
    {
        LB_CUDA(c, cudaStreamSynchronize(c->stream));
        if (dev_alloc(c, c->d_tkeys, table) || dev_alloc(c, c->d_tcount, table) || dev_alloc(c, c->d_tlive, table) || dev_alloc(c, c->d_cells, table))
            return LIDAR_B200_ERR_CUDA;
        c->cap_table = static_cast<uint32_t>(table);
    }
    const size_t planes = static_cast<size_t>(c->cap_frames) * c->seg.partitions * c->seg.iterations * 4;
    if (dev_alloc(c, c->d_planes, planes ? planes : 4) ||
        dev_alloc(c, c->d_status, static_cast<size_t>(c->cap_frames) * (c->seg.partitions ? c->seg.partitions : 1)))
        return LIDAR_B200_ERR_CUDA;
    if (dev_alloc(c, c->d_err, 4) || pin_alloc(c, c->h_err, 4))
        return LIDAR_B200_ERR_CUDA;
    return 0;
}

int apply_seg_cfg(lidar_b200_ctx *c, const lidar_b200_seg_cfg &cfg)
{
    if (cfg.number_of_planar_partitions == 0u || cfg.number_of_planar_partitions > 4096u)
        return fail(c, LIDAR_B200_ERR_UNSUPPORTED, "number_of_planar_partitions must be in [1, 4096]");
    if (cfg.number_of_iterations == 0u || cfg.number_of_iterations > 64u)
        return fail(c, LIDAR_B200_ERR_UNSUPPORTED, "number_of_iterations must be in [1, 64]");
    if (cfg.number_of_lower_point_representatives == 0u)
        return fail(c, LIDAR_B200_ERR_UNSUPPORTED, "number_of_lower_point_representatives must be at least 1");
    if (!std::isfinite(cfg.sensor_height_m) || !std::isfinite(cfg.orthogonal_distance_threshold) ||
        !std::isfinite(cfg.initial_seed_threshold))
        return fail(c, LIDAR_B200_ERR_INVALID, "non-finite segmentation configuration");
    c->seg_cfg = cfg;
    c->seg.sensor_height_m = cfg.sensor_height_m;
    c->seg.orthogonal_distance_threshold = cfg.orthogonal_distance_threshold;
    c->seg.initial_seed_threshold = cfg.initial_seed_threshold;
    c->seg.iterations = cfg.number_of_iterations;
    c->seg.partitions = cfg.number_of_planar_partitions;
    c->seg.lpr = cfg.number_of_lower_point_representatives;
    return 0;
}

int apply_clu_cfg(lidar_b200_ctx *c, const lidar_b200_clu_cfg &cfg)
{
    if (!(cfg.distance_squared > 0.0f) || !std::isfinite(cfg.distance_squared) || !std::isfinite(cfg.cluster_quality))
        return fail(c, LIDAR_B200_ERR_INVALID, "distance_squared must be finite and > 0, cluster_quality finite");
    c->clu_cfg = cfg;
    c->clu.distance_squared = cfg.distance_squared;
    // clustering.cpp:66-67: std::pow(1.0 - quality, 2) * distance_squared in double; `dist <= thr`
    // promotes the float d2. Equivalent float constant: the largest float not above thr.
    const double thr = std::pow(1.0 - static_cast<double>(cfg.cluster_quality), 2) * static_cast<double>(cfg.distance_squared);
    float tf = static_cast<float>(thr);
    if (static_cast<double>(tf) > thr)
        tf = std::nextafterf(tf, -std::numeric_limits<float>::infinity());
    c->clu.inner_threshold = tf;
    c->clu.min_cluster_size = cfg.min_cluster_size;
    c->clu.max_cluster_size = cfg.max_cluster_size;
    c->clu.inv_cell = 1.0 / (std::sqrt(static_cast<double>(cfg.distance_squared)) * 1.001);
    c->clu.cta_min_members = 256u;
    if (const char *e = std::getenv("LIDAR_B200_CTA_MIN_MEMBERS"))
    {
        const long v = std::atol(e);
        if (v >= static_cast<long>(kCtaComponentMin))
            c->clu.cta_min_members = static_cast<uint32_t>(v);
    }
    return 0;
}

// stage boundaries: 0 begin | 1 x-sort | 2 gather+fit | 3 compact | 4 voxel grid | 5 union-find |
// 6 component sort | 7 k-d order | 8 replay | 9 label compaction
int mark(lidar_b200_ctx *c, int idx)
{
    if (!c->profiling)
        return 0;
    LB_CUDA(c, cudaEventRecord(c->ev_stage[idx], c->stream));
    c->n_stage_marks = idx + 1;
    return 0;
}

uint32_t grid_x(uint32_t n, uint32_t per_block, uint32_t cap)
{
    uint32_t g = (n + per_block - 1u) / per_block;
    g = g < 1u ? 1u : g;
    return g > cap ? cap : g;
}

int finish_fetch(lidar_b200_ctx *c);
int enqueue_fetch_payload(lidar_b200_ctx *c, bool block);

// lays the frames of a batch out, stages the points into pinned memory and starts the upload
int stage(lidar_b200_ctx *c, uint32_t n_frames, const void *const *points, const uint32_t *n_points,
          uint32_t stride_bytes)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    if (stride_bytes < 12u || (stride_bytes & 3u))
        return fail(c, LIDAR_B200_ERR_INVALID, "stride_bytes must be a multiple of 4 and >= 12");
    LB_CUDA(c, cudaSetDevice(c->device));
    if (c->fetch.pending) // complete the batch this context still owes before its bookkeeping is reused
    {
        const int rc = finish_fetch(c);
        if (rc)
            return rc;
    }
    uint64_t total = 0;
    uint32_t max_n = 0;
    c->off.assign(n_frames, 0u);
    c->cnt.assign(n_frames, 0u);
    for (uint32_t f = 0; f < n_frames; ++f)
    {
        c->off[f] = static_cast<uint32_t>(total);
        c->cnt[f] = n_points[f];
        max_n = n_points[f] > max_n ? n_points[f] : max_n;
        total += (static_cast<uint64_t>(n_points[f]) + 31u) & ~31ull; // 128-byte aligned frame starts
        if (n_points[f] && !points[f])
            return fail(c, LIDAR_B200_ERR_INVALID, "null points pointer");
    }
    if (total > 0x7FFFFFFFull)
        return fail(c, LIDAR_B200_ERR_CAPACITY, "batch exceeds 2^31 points");
    uint64_t geom_total = total;
    if (n_frames == 1u && c->use_graph && max_n)
    {
        // one frame in flight: launch geometry in steps of kGraphBucket points, so that the captured graph of a frame
        // serves the next frames too (grids, tile counts and the sizes of the path's memsets / copies follow max_n and
        // geom_total; the kernels themselves stop at the frame's own count, which lives in device memory). The slot
        // layout the callers see (c->total, offsets) stays exact.
        max_n = (max_n + kGraphBucket - 1u) / kGraphBucket * kGraphBucket;
        geom_total = max_n;
    }
    if (geom_total > c->cap_pts || n_frames > c->cap_frames)
    {
        const int rc = reserve(c, geom_total > c->cap_pts ? static_cast<uint32_t>(geom_total) : c->cap_pts,
                               n_frames > c->cap_frames ? n_frames : c->cap_frames);
        if (rc)
            return rc;
    }
    // the previous batch may still be reading the staging buffers
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->n_frames = n_frames;
    c->total = static_cast<uint32_t>(total);
    c->geom_total = static_cast<uint32_t>(geom_total);
    c->max_n = max_n;
    uint32_t *hm = c->h_meta.p;
    const size_t F = c->cap_frames;
    uint32_t toff = 0, max_tcap = 0;
    for (uint32_t f = 0; f < n_frames; ++f)
    {
        hm[f] = c->off[f];
        hm[F + f] = c->cnt[f];
        const uint32_t tcap = next_pow2(c->cnt[f] * 2u < 64u ? 64u : c->cnt[f] * 2u);
        hm[2 * F + f] = toff;
        hm[3 * F + f] = tcap;
        toff += tcap;
        max_tcap = tcap > max_tcap ? tcap : max_tcap;
    }
    c->max_tcap = max_tcap;
    if (dev_alloc(c, c->d_hist, radix_sort_scratch_words(n_frames, max_n)))
        return LIDAR_B200_ERR_CUDA;
    if (n_frames)
        LB_CUDA(c, cudaMemcpyAsync(c->d_meta.p, hm, 4 * F * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    // one upload per frame, issued as soon as the frame is staged: the DMA of frame f overlaps the
    // staging pass of the frames behind it. Page-locked 16-byte records need no staging pass at all.
    // (asked on this thread: the staging threads below make no CUDA call)
    std::vector<uint8_t> direct(n_frames, 0u);
    for (uint32_t f = 0; f < n_frames; ++f)
        direct[f] = (c->cnt[f] == 0u || (stride_bytes == 16u && is_pinned(points[f]))) ? 1u : 0u;
    auto stage_frame = [&](uint32_t f) -> const void * {
        const uint32_t n = c->cnt[f];
        const uint8_t *src = static_cast<const uint8_t *>(points[f]);
        if (direct[f])
            return src;
        float4 *dst = c->h_pts.p + c->off[f];
        if (stride_bytes == 16u)
            std::memcpy(dst, src, static_cast<size_t>(n) * 16u);
        else if (stride_bytes >= 16u)
            for (uint32_t i = 0; i < n; ++i)
                std::memcpy(&dst[i], src + static_cast<size_t>(i) * stride_bytes, 16u);
        else
            for (uint32_t i = 0; i < n; ++i)
            {
                std::memcpy(&dst[i], src + static_cast<size_t>(i) * stride_bytes, 12u);
                dst[i].w = 1.0f;
            }
        return dst;
    };
    // A batch of pageable clouds is staged by a few threads (frames handed out through a counter, uploads still issued
    // in frame order by this thread): one thread copies ~15 GB/s, which capped the pipeline at ~4 000 frames/s.
    uint32_t n_workers = 0u;
    if (n_frames >= 4u && total >= (1u << 20) && !direct[0])
    {
        n_workers = c->stage_threads;
        if (n_workers > n_frames / 2u)
            n_workers = n_frames / 2u;
    }
    if (n_workers <= 1u)
    {
        for (uint32_t f = 0; f < n_frames; ++f)
        {
            if (c->cnt[f] == 0u)
                continue;
            const void *from = stage_frame(f);
            LB_CUDA(c, cudaMemcpyAsync(c->d_pts.p + c->off[f], from, static_cast<size_t>(c->cnt[f]) * sizeof(float4),
                                       cudaMemcpyHostToDevice, c->stream));
        }
        return 0;
    }
    std::vector<const void *> from(n_frames, nullptr);
    std::vector<std::atomic<uint32_t>> done(n_frames);
    for (auto &d : done)
        d.store(0u, std::memory_order_relaxed);
    std::atomic<uint32_t> next{0u};
    std::vector<std::thread> workers;
    workers.reserve(n_workers);
    const auto stage_loop = [&]() {
        for (uint32_t f = next.fetch_add(1u); f < n_frames; f = next.fetch_add(1u))
        {
            from[f] = stage_frame(f);
            done[f].store(1u, std::memory_order_release);
        }
    };
    for (uint32_t t = 0; t < n_workers; ++t)
    {
        try
        {
            workers.emplace_back(stage_loop);
        }
        catch (...) // no more threads to be had: the ones that started (or this one, below) do the work
        {
            break;
        }
    }
    if (workers.empty())
        stage_loop();
    cudaError_t first_error = cudaSuccess;
    for (uint32_t f = 0; f < n_frames; ++f)
    {
        while (done[f].load(std::memory_order_acquire) == 0u)
            std::this_thread::yield();
        if (c->cnt[f] == 0u || first_error != cudaSuccess)
            continue;
        first_error = cudaMemcpyAsync(c->d_pts.p + c->off[f], from[f], static_cast<size_t>(c->cnt[f]) * sizeof(float4),
                                      cudaMemcpyHostToDevice, c->stream);
    }
    for (auto &w : workers)
        w.join();
    LB_CUDA(c, first_error);
    return 0;
}

int run_segmentation(lidar_b200_ctx *c)
{
    const uint32_t F = c->n_frames;
    if (F == 0u)
        return 0;
    cudaStream_t s = c->stream;
    LB_CUDA(c, cudaMemsetAsync(c->d_labels.p, 0, static_cast<size_t>(c->geom_total ? c->geom_total : 1) * 4, s));
    LB_CUDA(c, cudaMemsetAsync(c->m_ng(), 0, 2 * static_cast<size_t>(c->cap_frames) * 4, s)); // n_ground, n_obstacle
    if (c->max_n == 0u)
        return 0;
    const BatchView bv{c->m_off(), c->m_cnt(), F};
    const dim3 gp(grid_x(c->max_n, 256u, 2048u), F);
    mark(c, 0);
    seg_keys_kernel<<<gp, 256, 0, s>>>(c->d_pts.p, bv, c->d_key_a.p, c->d_val_a.p);
    ++c->launches;
    int rl = 0;
    // (the frame-resident sort below was tried here too, with the key generation and the gather fused into its first and
    // last pass: 1.9 ms against 1.35 + 0.15 ms per 154 frames - one CTA per frame leaves the SMs at 16 warps)
    const int passes = radix_sort_pairs(s, c->d_key_a.p, c->d_val_a.p, c->d_key_b.p, c->d_val_b.p, bv, c->max_n, 32u,
                                        RadixSortScratch{c->d_hist.p}, &rl);
    c->launches += rl;
    const uint32_t *sorted_idx = (passes & 1) ? c->d_val_b.p : c->d_val_a.p;
    mark(c, 1);
    uint32_t *zkeys = c->d_key_a.p; // the sort's key buffers are free again; the order lives in val_a / val_b
    seg_gather_kernel<<<gp, 256, 0, s>>>(c->d_pts.p, sorted_idx, bv, c->d_spts.p, zkeys);
    ++c->launches;
    seg_fit_kernel<<<dim3(c->seg.partitions, F), kFitThreads, sizeof(FitSmem), s>>>(c->d_spts.p, zkeys, bv, c->seg, c->d_flags.p,
                                                                                    c->d_planes.p, c->d_status.p,
                                                                                    c->d_key_b.p /* spare since the x sort */);
    ++c->launches;
    mark(c, 2);
    {
        const uint32_t tiles = grid_x(c->max_n, kCompactTile, 0xFFFFu);
        unsigned long long *tile_counts = reinterpret_cast<unsigned long long *>(c->d_hist.p);
        seg_compact_count_kernel<<<dim3(tiles, F), 1024, 0, s>>>(c->d_flags.p, bv, tiles, tile_counts);
        seg_compact_kernel<<<dim3(tiles, F), 1024, 0, s>>>(c->d_spts.p, c->d_flags.p, bv, tiles, tile_counts, c->d_labels.p,
                                                           c->d_gidx.p, c->d_oidx.p, c->d_obs.p, c->m_ng(), c->m_no());
    }
    c->launches += 2;
    mark(c, 3);
    LB_CUDA(c, cudaGetLastError());
    return 0;
}

// pts: cloud to cluster (frame-major, same offsets); counts: device array of per-frame sizes
int run_clustering(lidar_b200_ctx *c, const float4 *pts, const uint32_t *counts, uint32_t max_m)
{
    const uint32_t F = c->n_frames;
    if (F == 0u)
        return 0;
    cudaStream_t s = c->stream;
    LB_CUDA(c, cudaMemsetAsync(c->m_nc(), 0, static_cast<size_t>(c->cap_frames) * 4, s));
    LB_CUDA(c, cudaMemsetAsync(c->d_err.p, 0, 4, s));
    c->clu_pts = pts;
    c->clu_counts = counts;
    c->clu_max_m = max_m;
    c->grouped = false;
    c->hulled = false;
    if (max_m == 0u)
        return 0;
    const BatchView bv{c->m_off(), counts, F};
    const TableView tv{c->m_toff(), c->m_tcap()};
    const dim3 gp(grid_x(max_m, 256u, 2048u), F);
    const dim3 gt(grid_x(c->max_tcap, 256u, 2048u), F);
    if (c->batch_is_cluster_only)
        mark(c, 3);

    // The k-d pre-order ranks depend on the cloud only: they are built on the second stream while this
    // one builds the voxel grid and the components (stage 'kd_order' below is what is left to wait for).
    LB_CUDA(c, cudaEventRecord(c->ev_fork, s));
    LB_CUDA(c, cudaStreamWaitEvent(c->stream_big, c->ev_fork, 0));
    c->launches += kd_build_launch(c->stream_big, pts, bv, max_m, c->d_nodes.p, c->d_gepos.p, c->d_lepos.p, c->d_rank.p);
    LB_CUDA(c, cudaEventRecord(c->ev_kd, c->stream_big));

    grid_clear_kernel<<<gt, 256, 0, s>>>(bv, tv, c->d_tkeys.p, c->d_tcount.p);
    grid_insert_kernel<<<gp, 256, 0, s>>>(pts, bv, tv, c->clu, c->d_tkeys.p, c->d_tcount.p, c->d_slot_of.p, c->d_err.p);
    {
        const uint32_t tiles = grid_x(c->max_tcap, kScanTile, 0xFFFFu);
        unsigned long long *tile_counts = reinterpret_cast<unsigned long long *>(c->d_hist.p);
        grid_scan_count_kernel<<<dim3(tiles, F), 1024, 0, s>>>(bv, tv, c->d_tcount.p, tiles, tile_counts);
        grid_scan_kernel<<<dim3(tiles, F), 1024, 0, s>>>(bv, tv, c->d_tkeys.p, c->d_tcount.p, c->d_cells.p, tiles, tile_counts);
    }
    grid_fill_kernel<<<gp, 256, 0, s>>>(pts, bv, tv, c->d_cells.p, c->d_tcount.p, c->d_slot_of.p, c->d_cpts.p,
                                        c->d_pos_of.p, c->d_state.p /* cell_of */);
    mark(c, 4);
    cc_init_kernel<<<gp, 256, 0, s>>>(bv, c->d_parent.p, c->d_comp_size.p);
    {
        const uint32_t *cell_of = c->d_state.p;
        cc_nbr_kernel<<<dim3(grid_x(max_m, 256u, 2048u), F), 256, 0, s>>>(c->d_cpts.p, bv, tv, c->d_cells.p, c->d_slot_of.p,
                                                                          cell_of, c->d_nbr.p, c->d_cinfo.p, c->d_nb27.p);
        cc_sample_kernel<<<gp, 256, 0, s>>>(c->d_cpts.p, bv, c->clu, cell_of, c->d_nbr.p, c->d_cinfo.p, c->d_parent.p);
        cc_compress_kernel<<<gp, 256, 0, s>>>(bv, c->d_parent.p);
        cc_cell_parent_kernel<<<gp, 256, 0, s>>>(bv, cell_of, c->d_parent.p, c->d_cinfo.p);
        cc_link_kernel<<<gp, 256, 0, s>>>(c->d_cpts.p, bv, c->clu, cell_of, c->d_nbr.p, c->d_cinfo.p, c->d_parent.p);
    }
    cc_flatten_kernel<<<gp, 256, 0, s>>>(bv, c->d_parent.p, c->d_pos_of.p, c->d_key_a.p, c->d_val_a.p, c->d_comp_size.p);
    c->launches += 12;
    mark(c, 5);
    LB_CUDA(c, cudaMemcpyAsync(c->d_root.p, c->d_key_a.p, static_cast<size_t>(c->geom_total) * 4, cudaMemcpyDeviceToDevice, s));
    int rl = 0;
    const uint32_t bits = ceil_log2(max_m < 2u ? 2u : max_m);
    int passes;
    if (c->frame_sort && F >= kFsMinFrames)
    {
        passes = static_cast<int>((bits + 7u) / 8u);
        LB_CUDA(c, rs_frame_sort_launch(s, c->sm_count, c->d_key_a.p, c->d_val_a.p, c->d_key_b.p, c->d_val_b.p, bv, passes));
        rl = 1;
    }
    else
        passes = radix_sort_pairs(s, c->d_key_a.p, c->d_val_a.p, c->d_key_b.p, c->d_val_b.p, bv, max_m, bits,
                                  RadixSortScratch{c->d_hist.p}, &rl);
    c->launches += rl;
    const uint32_t *member_root = (passes & 1) ? c->d_key_b.p : c->d_key_a.p;
    const uint32_t *member_idx = (passes & 1) ? c->d_val_b.p : c->d_val_a.p;

    mark(c, 6);
    LB_CUDA(c, cudaStreamWaitEvent(s, c->ev_kd, 0)); // k-d order built on the second stream meanwhile

    mark(c, 7);
    {
    replay_init_kernel<<<gp, 256, 0, s>>>(c->d_cpts.p, bv, c->d_rank.p, member_idx, c->d_pos_of.p, c->d_slot_of.p,
                                          c->d_rpts.p, c->d_seed_of.p, c->d_member_pos.p, c->d_pslot.p, c->m_cursor(), tv,
                                          c->d_cells.p, c->d_pkey.p);
    replay_live_init_kernel<<<gt, 256, 0, s>>>(bv, tv, c->d_cells.p, c->d_tlive.p);
    // components of at least kCtaComponent members: one CTA each (speculative rounds, CTA-wide scans),
    // started first on a second stream so that the longest BFS chains begin at time zero
    uint32_t *big_count = c->m_cursor() + 3; // kBigBuckets counters
    const uint32_t bucket_capacity = c->cap_pts / kCtaComponentMin + 1u;
    bool huge_launched = false, small_launched = false;
    if (c->replay_version >= 5 && max_m < (1u << kGenPosBits))
    {
        // fifth generation (replay_gen.cuh): window-synchronous replay, three state planes in shared memory. The long jobs
        // (components of 2048 members and more: their window chains bound the launch) get CTAs of `replay5_big_threads`
        // threads, the short ones share SMs with 256-thread CTAs on a stream of their own. Components beyond the normal
        // planes go to a 1-CTA-per-SM launch with large planes, beyond that to the first generation.
        const uint32_t normal_words = 1024u; // 32 768 members
        const uint32_t huge_words = static_cast<uint32_t>((227u * 1024u - sizeof(GenSmem)) / 12u) & ~31u;
        const size_t smem_normal = sizeof(GenSmem) + 12u * normal_words, smem_huge = sizeof(GenSmem) + 12u * huge_words;
        if (!c->replay5_attr_done)
        {
#define LB_GEN_ATTR(NT, MINB, BYTES)                                                                                  \
    LB_CUDA(c, cudaFuncSetAttribute(replay_gen_kernel<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(BYTES)))
            LB_GEN_ATTR(256, 2, smem_normal);
            LB_GEN_ATTR(256, 3, smem_normal);
            LB_GEN_ATTR(512, 2, smem_normal);
            LB_GEN_ATTR(1024, 1, smem_huge);
#undef LB_GEN_ATTR
            c->replay5_attr_done = true;
        }
        uint32_t *mcell = reinterpret_cast<uint32_t *>(c->d_mkey.p);
        replay_init5_kernel<<<gp, 256, 0, s>>>(c->d_cpts.p, bv, c->d_rank.p, member_idx, c->d_pos_of.p, c->d_state.p /* cell_of */,
                                               c->d_ipts.p, c->d_rankpos.p, c->d_mpts.p, mcell);
        replay_biglist2_kernel<<<gp, 256, 0, s>>>(bv, member_root, c->d_comp_size.p, c->clu.cta_min_members, 32u * normal_words,
                                                  32u * huge_words, c->d_biglist.p, bucket_capacity, big_count,
                                                  c->m_cursor() + 8, c->m_cursor() + 10);
        c->launches += 2;
        LB_CUDA(c, cudaEventRecord(c->ev_fork, s));
        LB_CUDA(c, cudaStreamWaitEvent(c->stream_big, c->ev_fork, 0));
#define LB_GEN_ARGS(LIST, COUNT, NB, CURSOR, WORDS, STATS, SKIP)                                                      \
    c->d_ipts.p, c->d_rankpos.p, c->d_mpts.p, mcell, c->d_nb27.p, c->d_cinfo.p, bv, c->clu, member_root, member_idx,  \
        c->d_comp_size.p, c->d_seed_of.p, c->d_queue.p, c->d_spill.p, c->d_seed_valid.p, LIST, bucket_capacity, COUNT, NB, \
        CURSOR, WORDS, STATS, SKIP
        const bool three_per_sm = c->replay5_ctas_per_sm >= 3u;
        const uint32_t big_threads = c->replay5_big_threads ? c->replay5_big_threads : (F <= 4u ? 512u : 256u);
        if (big_threads <= 256u)
        {
            // one launch for every job list
            if (three_per_sm)
                replay_gen_kernel<256, 3><<<c->sm_count * 3u, 256, smem_normal, c->stream_big>>>(
                    LB_GEN_ARGS(c->d_biglist.p, big_count, kBigBuckets, c->m_cursor() + 2, normal_words, c->d_job_stats.p, 0u));
            else
                replay_gen_kernel<256, 2><<<c->sm_count * 2u, 256, smem_normal, c->stream_big>>>(
                    LB_GEN_ARGS(c->d_biglist.p, big_count, kBigBuckets, c->m_cursor() + 2, normal_words, c->d_job_stats.p, 0u));
            ++c->launches;
        }
        else
        {
            // the first `wl` job lists (list 0: >= 8192 members, list 1: >= 2048) first, with wide CTAs; the others beside them
            const uint32_t wl = c->replay5_wide_lists;
            const uint32_t wide_grid = c->sm_count * c->replay5_wide_ctas_per_sm;
            if (big_threads >= 1024u)
                replay_gen_kernel<1024, 1><<<c->sm_count, 1024, smem_normal, c->stream_big>>>(
                    LB_GEN_ARGS(c->d_biglist.p, big_count, wl, c->m_cursor() + 2, normal_words, c->d_job_stats.p, 0u));
            else
                replay_gen_kernel<512, 2><<<wide_grid, 512, smem_normal, c->stream_big>>>(
                    LB_GEN_ARGS(c->d_biglist.p, big_count, wl, c->m_cursor() + 2, normal_words, c->d_job_stats.p, 0u));
            LB_CUDA(c, cudaStreamWaitEvent(c->stream_small, c->ev_fork, 0));
            if (three_per_sm)
                replay_gen_kernel<256, 3><<<c->sm_count * 3u, 256, smem_normal, c->stream_small>>>(
                    LB_GEN_ARGS(c->d_biglist.p + wl * static_cast<size_t>(bucket_capacity), big_count + wl, kBigBuckets - wl,
                                c->m_cursor() + 14, normal_words, c->d_job_stats.p, wl));
            else
                replay_gen_kernel<256, 2><<<c->sm_count * 2u, 256, smem_normal, c->stream_small>>>(
                    LB_GEN_ARGS(c->d_biglist.p + wl * static_cast<size_t>(bucket_capacity), big_count + wl, kBigBuckets - wl,
                                c->m_cursor() + 14, normal_words, c->d_job_stats.p, wl));
            LB_CUDA(c, cudaEventRecord(c->ev_join3, c->stream_small));
            small_launched = true;
            c->launches += 2;
        }
        if (max_m > 32u * normal_words) // a component can only be that large in a frame that large
        {
            LB_CUDA(c, cudaStreamWaitEvent(c->stream_huge, c->ev_fork, 0));
            replay_gen_kernel<1024, 1><<<c->sm_count, 1024, smem_huge, c->stream_huge>>>(
                LB_GEN_ARGS(c->d_biglist.p + 4u * static_cast<size_t>(bucket_capacity), c->m_cursor() + 8, 1u, c->m_cursor() + 7,
                            huge_words, nullptr, 0u));
            ++c->launches;
            if (max_m > 32u * huge_words)
            {
                // first generation for what is left: its list is bucket 5, its counters sit at cursor[10..13] (11..13 stay 0)
                replay_cta_kernel<3><<<c->sm_count * 3u, kCtaThreads, 0, c->stream_huge>>>(
                    c->d_rpts.p, bv, tv, c->d_cells.p, c->clu, member_root, member_idx, c->d_member_pos.p, c->d_comp_size.p,
                    c->d_pkey.p, c->d_tlive.p, c->d_seed_of.p, c->d_queue.p, c->d_spill.p, c->d_seed_valid.p,
                    c->d_biglist.p + 5u * static_cast<size_t>(bucket_capacity), bucket_capacity, c->m_cursor() + 10,
                    c->m_cursor() + 9, nullptr);
                ++c->launches;
            }
            LB_CUDA(c, cudaEventRecord(c->ev_join2, c->stream_huge));
            huge_launched = true;
        }
#undef LB_GEN_ARGS
    }
    else if (c->replay_version >= 2)
    {
        // second generation (replay_cta2.cuh): component state in shared memory, immutable point records. Components
        // beyond the normal bitmap go to a 1-CTA-per-SM launch with a large bitmap, beyond that to the first generation.
        const uint32_t normal_words = kReplay2NormalWords, huge_words = kReplay2HugeWords;
        const size_t smem_normal = sizeof(Cta2Smem) + 4u * normal_words, smem_huge = sizeof(Cta2Smem) + 4u * huge_words;
        if (!c->replay2_attr_done)
        {
            LB_CUDA(c, cudaFuncSetAttribute(replay_cta2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_normal)));
            LB_CUDA(c, cudaFuncSetAttribute(replay_cta2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_huge)));
            c->replay2_attr_done = true;
        }
        replay_init2_kernel<<<gp, 256, 0, s>>>(c->d_cpts.p, bv, c->d_rank.p, member_idx, c->d_pos_of.p, c->d_pkey.p, c->d_ipts.p,
                                               c->d_rankpos.p, c->d_mpts.p, c->d_mkey.p);
        replay_biglist2_kernel<<<gp, 256, 0, s>>>(bv, member_root, c->d_comp_size.p, c->clu.cta_min_members, 16u * normal_words,
                                                  16u * huge_words, c->d_biglist.p, bucket_capacity, big_count,
                                                  c->m_cursor() + 8, c->m_cursor() + 10);
        c->launches += 2;
        LB_CUDA(c, cudaEventRecord(c->ev_fork, s));
        LB_CUDA(c, cudaStreamWaitEvent(c->stream_big, c->ev_fork, 0));
        replay_cta2_kernel<3><<<c->sm_count * 3u, kCtaThreads, smem_normal, c->stream_big>>>(
            c->d_ipts.p, c->d_rankpos.p, c->d_mpts.p, c->d_mkey.p, bv, tv, c->d_cells.p, c->clu, member_root, member_idx,
            c->d_comp_size.p, c->d_tlive.p, c->d_seed_of.p, c->d_queue.p, c->d_spill.p, c->d_seed_valid.p, c->d_biglist.p,
            bucket_capacity, big_count, kBigBuckets, c->m_cursor() + 2, normal_words, c->d_job_stats.p);
        ++c->launches;
        if (max_m > 16u * normal_words) // a component can only be that large in a frame that large
        {
            LB_CUDA(c, cudaStreamWaitEvent(c->stream_huge, c->ev_fork, 0));
            replay_cta2_kernel<1><<<c->sm_count, kCtaThreads, smem_huge, c->stream_huge>>>(
                c->d_ipts.p, c->d_rankpos.p, c->d_mpts.p, c->d_mkey.p, bv, tv, c->d_cells.p, c->clu, member_root, member_idx,
                c->d_comp_size.p, c->d_tlive.p, c->d_seed_of.p, c->d_queue.p, c->d_spill.p, c->d_seed_valid.p,
                c->d_biglist.p + 4u * static_cast<size_t>(bucket_capacity), bucket_capacity, c->m_cursor() + 8, 1u,
                c->m_cursor() + 7, huge_words, nullptr);
            ++c->launches;
            if (max_m > 16u * huge_words)
            {
                // first generation for what is left: its list is bucket 5, its counters sit at cursor[10..13] (11..13 stay 0)
                replay_cta_kernel<3><<<c->sm_count * 3u, kCtaThreads, 0, c->stream_huge>>>(
                    c->d_rpts.p, bv, tv, c->d_cells.p, c->clu, member_root, member_idx, c->d_member_pos.p, c->d_comp_size.p,
                    c->d_pkey.p, c->d_tlive.p, c->d_seed_of.p, c->d_queue.p, c->d_spill.p, c->d_seed_valid.p,
                    c->d_biglist.p + 5u * static_cast<size_t>(bucket_capacity), bucket_capacity, c->m_cursor() + 10,
                    c->m_cursor() + 9, nullptr);
                ++c->launches;
            }
            LB_CUDA(c, cudaEventRecord(c->ev_join2, c->stream_huge));
            huge_launched = true;
        }
    }
    else
    {
    replay_biglist_kernel<<<gp, 256, 0, s>>>(bv, member_root, c->d_comp_size.p, c->clu.cta_min_members, c->d_biglist.p,
                                             bucket_capacity, big_count);
    LB_CUDA(c, cudaEventRecord(c->ev_fork, s));
    LB_CUDA(c, cudaStreamWaitEvent(c->stream_big, c->ev_fork, 0));
    {
        const uint32_t per_sm = c->replay_big_ctas_per_sm;
        const uint32_t grid = c->sm_count * per_sm;
#define LB_LAUNCH_REPLAY_CTA(MINB)                                                                                     \
    replay_cta_kernel<MINB><<<grid, kCtaThreads, 0, c->stream_big>>>(                                                  \
        c->d_rpts.p, bv, tv, c->d_cells.p, c->clu, member_root, member_idx, c->d_member_pos.p, c->d_comp_size.p,       \
        c->d_pkey.p, c->d_tlive.p, c->d_seed_of.p, c->d_queue.p, c->d_spill.p, c->d_seed_valid.p, c->d_biglist.p,      \
        bucket_capacity, big_count, c->m_cursor() + 2, c->d_job_stats.p)
        if (per_sm >= 4u)
            LB_LAUNCH_REPLAY_CTA(4);
        else if (per_sm == 3u)
            LB_LAUNCH_REPLAY_CTA(3);
        else
            LB_LAUNCH_REPLAY_CTA(2);
#undef LB_LAUNCH_REPLAY_CTA
    }
    c->launches += 2;
    }
    LB_CUDA(c, cudaEventRecord(c->ev_join, c->stream_big));
    const uint32_t claims = (max_m + 31u) / 32u;
    // persistent grid: a fixed number of CTAs per SM walks the flat (frame, claim) work list
    const uint32_t rctas = grid_x(claims * F, kReplayWarps, c->sm_count * c->replay_ctas_per_sm);
    replay_kernel<<<rctas, kReplayWarps * 32, 0, s>>>(c->d_rpts.p, bv, tv, c->d_cells.p, c->clu, member_root, member_idx,
                                                      c->d_member_pos.p, c->d_comp_size.p, c->d_pslot.p, c->d_tlive.p,
                                                      c->d_seed_of.p, c->d_queue.p, c->d_spill.p, c->d_seed_valid.p,
                                                      c->m_cursor(), claims);
    LB_CUDA(c, cudaStreamWaitEvent(s, c->ev_join, 0));
    if (huge_launched)
        LB_CUDA(c, cudaStreamWaitEvent(s, c->ev_join2, 0));
    if (small_launched)
        LB_CUDA(c, cudaStreamWaitEvent(s, c->ev_join3, 0));
    }
    mark(c, 8);
    {
        const uint32_t tiles = grid_x(max_m, kLabelTile, 0xFFFFu);
        unsigned long long *tile_counts = reinterpret_cast<unsigned long long *>(c->d_hist.p);
        label_count_kernel<<<dim3(tiles, F), 1024, 0, s>>>(bv, c->d_pos_of.p, c->d_seed_of.p, c->d_seed_valid.p, tiles,
                                                           tile_counts);
        label_number_kernel<<<dim3(tiles, F), 1024, 0, s>>>(bv, c->d_pos_of.p, c->d_seed_of.p, c->d_seed_valid.p, tiles,
                                                            tile_counts, c->d_seed_label.p, c->m_nc());
        label_assign_kernel<<<gp, 256, 0, s>>>(bv, c->d_pos_of.p, c->d_seed_of.p, c->d_seed_valid.p, c->d_seed_label.p,
                                               c->d_clabels.p);
    }
    c->launches += 6;
    mark(c, 9);
    LB_CUDA(c, cudaGetLastError());
    return 0;
}

// completes an asynchronous fetch: waits for the stream, un-stages the results that could not be
// written by DMA directly and reports the per-frame counts
// Second phase of a fetch: exact-size copies of the result arrays. `block` = wait for the counts (else return
// without doing anything while the batch is still running).
int enqueue_fetch_payload(lidar_b200_ctx *c, bool block)
{
    lidar_b200_ctx::FetchReq &r = c->fetch;
    if (!r.pending || r.payload_enqueued)
        return 0;
    LB_CUDA(c, cudaSetDevice(c->device));
    if (block)
        LB_CUDA(c, cudaEventSynchronize(c->ev_counts));
    else
    {
        const cudaError_t q = cudaEventQuery(c->ev_counts);
        if (q == cudaErrorNotReady)
            return 0;
        if (q != cudaSuccess)
            return fail(c, LIDAR_B200_ERR_CUDA, std::string("cudaEventQuery: ") + cudaGetErrorString(q));
    }
    cudaStream_t s = c->stream;
    const size_t F = c->cap_frames;
    const uint32_t *hm = c->h_meta.p;
    const void *dev[4] = {c->d_labels.p, c->d_gidx.p, c->d_oidx.p, c->d_clabels.p};
    r.payload_enqueued = true; // (also when a copy below fails: the batch is lost either way)
    // copy list: the segmentation labels of all frames in one piece (the slots are contiguous, at most 31 padding
    // entries per frame), the three per-frame lists at their exact sizes
    std::vector<void *> &dsts = c->copy_dst, &srcs = c->copy_src;
    std::vector<size_t> &sizes = c->copy_size;
    dsts.clear();
    srcs.clear();
    sizes.clear();
    auto add = [&](int k, size_t first, size_t count) {
        if (!r.out[k] || count == 0u)
            return;
        dsts.push_back((r.direct[k] ? static_cast<uint32_t *>(r.out[k]) : c->h_u32[k].p) + first);
        srcs.push_back(const_cast<uint32_t *>(static_cast<const uint32_t *>(dev[k])) + first);
        sizes.push_back(count * 4u);
    };
    if (c->n_frames)
        add(0, c->off[0], static_cast<size_t>(c->off[c->n_frames - 1u]) + c->cnt[c->n_frames - 1u] - c->off[0]);
    if (c->fetch_mode == 1 && c->n_frames)
        for (int k = 1; k < 4; ++k)
            add(k, 0, c->total);
    for (uint32_t f = 0; f < c->n_frames && c->fetch_mode != 1; ++f)
    {
        // (counts of a batch that raised the input-error flag are not trusted beyond the slot)
        const uint32_t n_ground = hm[4 * F + f] < c->cnt[f] ? hm[4 * F + f] : c->cnt[f];
        const uint32_t n_obstacle = hm[5 * F + f] < c->cnt[f] ? hm[5 * F + f] : c->cnt[f];
        add(1, c->off[f], n_ground);
        add(2, c->off[f], n_obstacle);
        add(3, c->off[f], n_obstacle);
    }
    if (sizes.empty())
        return 0;
    // one driver call for the whole list (cudaMemcpyBatchAsync, CUDA >= 12.8); plain copies if it is refused
    static std::atomic<bool> batch_ok{true}; // (the pipeline's worker thread and the caller's thread both get here)
    if (batch_ok && sizes.size() > 1u && c->fetch_mode == 3)
    {
        cudaMemcpyAttributes attr{};
        attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        size_t attr_idx = 0, fail_idx = 0;
        const cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), sizes.size(), &attr, &attr_idx, 1,
                                                   &fail_idx, s);
        if (e == cudaSuccess)
            return 0;
        (void)cudaGetLastError();
        batch_ok = false;
    }
    for (size_t i = 0; i < sizes.size(); ++i)
        LB_CUDA(c, cudaMemcpyAsync(dsts[i], srcs[i], sizes[i], cudaMemcpyDeviceToHost, s));
    return 0;
}

int finish_fetch(lidar_b200_ctx *c)
{
    lidar_b200_ctx::FetchReq &r = c->fetch;
    {
        const int rc = enqueue_fetch_payload(c, true);
        if (rc)
        {
            r.pending = false;
            return rc;
        }
    }
    r.pending = false;
    LB_CUDA(c, cudaSetDevice(c->device));
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (cudaEventElapsedTime(&c->last_run_ms, c->ev_start, c->ev_stop) != cudaSuccess)
    {
        c->last_run_ms = 0.0f;
        (void)cudaGetLastError();
    }
    if (c->h_err.p[0])
        return fail(c, LIDAR_B200_ERR_INPUT, "non-finite or out-of-range point coordinates");
    const size_t F = c->cap_frames;
    const uint32_t *hm = c->h_meta.p;
    for (uint32_t f = 0; f < c->n_frames; ++f)
    {
        if (r.point_offset)
            r.point_offset[f] = c->off[f];
        if (r.n_ground)
            r.n_ground[f] = hm[4 * F + f];
        if (r.n_obstacle)
            r.n_obstacle[f] = hm[5 * F + f];
        if (r.n_clusters)
            r.n_clusters[f] = hm[6 * F + f];
    }
    for (uint32_t f = 0; f < c->n_frames; ++f)
    {
        const size_t o = c->off[f];
        const uint32_t n_ground = hm[4 * F + f] < c->cnt[f] ? hm[4 * F + f] : c->cnt[f];
        const uint32_t n_obstacle = hm[5 * F + f] < c->cnt[f] ? hm[5 * F + f] : c->cnt[f];
        const uint32_t used[4] = {c->cnt[f], n_ground, n_obstacle, n_obstacle};
        for (int k = 0; k < 4; ++k)
            if (r.out[k] && !r.direct[k] && used[k])
                std::memcpy(static_cast<uint32_t *>(r.out[k]) + o, c->h_u32[k].p + o, static_cast<size_t>(used[k]) * 4);
    }
    return 0;
}

} // namespace

namespace
{
void drop_graphs(lidar_b200_ctx *c)
{
    for (auto &g : c->graphs)
        if (g.exec)
            cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
}

int run_frames(lidar_b200_ctx *c)
{
    const int rc = run_segmentation(c);
    return rc ? rc : run_clustering(c, c->d_obs.p, c->m_no(), c->max_n);
}
} // namespace

extern "C"
{

void lidar_b200_seg_cfg_default(lidar_b200_seg_cfg *cfg)
{
    cfg->sensor_height_m = 1.73f;
    cfg->orthogonal_distance_threshold = 0.3f;
    cfg->initial_seed_threshold = 0.6f;
    cfg->number_of_iterations = 3u;
    cfg->number_of_planar_partitions = 2u;
    cfg->number_of_lower_point_representatives = 5000u;
}

void lidar_b200_clu_cfg_default(lidar_b200_clu_cfg *cfg)
{
    cfg->distance_squared = 0.18f;
    cfg->cluster_quality = 0.5f;
    cfg->min_cluster_size = 4u;
    cfg->max_cluster_size = 0xFFFFFFFFu;
}

int lidar_b200_create(int device, uint32_t max_points, uint32_t max_frames, lidar_b200_ctx **ctx_out)
{
    if (!ctx_out)
        return LIDAR_B200_ERR_INVALID;
    *ctx_out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return LIDAR_B200_ERR_CUDA;
    lidar_b200_ctx *c = new lidar_b200_ctx();
    c->device = device;
    c->want_job_stats = std::getenv("LIDAR_B200_REPLAY_STATS") != nullptr;
    c->chi_stats = std::getenv("LIDAR_B200_CHI_STATS") != nullptr;
    if (const char *e = std::getenv("LIDAR_B200_REPLAY5_WIDE_LISTS"))
        c->replay5_wide_lists = std::atoi(e) == 1 ? 1u : 2u;
    if (const char *e = std::getenv("LIDAR_B200_REPLAY5_WIDE_CTAS_PER_SM"))
        c->replay5_wide_ctas_per_sm = std::atoi(e) == 1 ? 1u : 2u;
    if (const char *e = std::getenv("LIDAR_B200_CHI_CTAS_PER_SM"))
        c->chi_ctas_per_sm = static_cast<uint32_t>(std::atoi(e) > 0 ? std::atoi(e) : 1);
    if (const char *e = std::getenv("LIDAR_B200_FETCH_MODE"))
        c->fetch_mode = std::atoi(e);
    if (const char *e = std::getenv("LIDAR_B200_GRAPH"))
        c->use_graph = std::atoi(e) != 0;
    if (const char *e = std::getenv("LIDAR_B200_STAGE_THREADS"))
        c->stage_threads = static_cast<uint32_t>(std::atoi(e) > 0 ? std::atoi(e) : 1);
    if (const char *e = std::getenv("LIDAR_B200_REPLAY_V"))
        c->replay_version = (std::atoi(e) == 1 || std::atoi(e) == 2) ? std::atoi(e) : 5;
    if (const char *e = std::getenv("LIDAR_B200_FRAME_SORT"))
        c->frame_sort = std::atoi(e) != 0;
    if (const char *e = std::getenv("LIDAR_B200_REPLAY5_BIG_THREADS"))
        c->replay5_big_threads = static_cast<uint32_t>(std::atoi(e));
    if (const char *e = std::getenv("LIDAR_B200_REPLAY5_CTAS_PER_SM"))
        c->replay5_ctas_per_sm = static_cast<uint32_t>(std::max(2, std::min(4, std::atoi(e))));
    lidar_b200_seg_cfg sc;
    lidar_b200_clu_cfg cc;
    lidar_b200_seg_cfg_default(&sc);
    lidar_b200_clu_cfg_default(&cc);
    apply_seg_cfg(c, sc);
    apply_clu_cfg(c, cc);
    // (a high-priority second stream - the 17 k-d levels and the long replay jobs first - was measured: the k-d order is
    // ready in time then, but the union-find beside it slows down by the same amount; a step is bound by the total work)
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream_big, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream_huge, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->stream_small, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join3, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&c->ev_start) != cudaSuccess || cudaEventCreate(&c->ev_stop) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_kd, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_counts, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess || // the waiter sleeps
       
        [&]() {
            for (auto &e : c->ev_stage)
                if (cudaEventCreate(&e) != cudaSuccess)
                    return true;
            return false;
        }() ||
        cudaFuncSetAttribute(seg_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(sizeof(FitSmem))) != cudaSuccess ||
        reserve(c, max_points ? max_points : 200000u, max_frames ? max_frames : 1u) != 0)
    {
        lidar_b200_destroy(c);
        return LIDAR_B200_ERR_CUDA;
    }
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0)
            c->sm_count = static_cast<uint32_t>(sms);
        if (const char *e = std::getenv("LIDAR_B200_REPLAY_CTAS_PER_SM"))
        {
            const int v = std::atoi(e);
            if (v >= 1 && v <= 16)
                c->replay_ctas_per_sm = static_cast<uint32_t>(v);
        }
        if (const char *e = std::getenv("LIDAR_B200_REPLAY_BIG_CTAS_PER_SM"))
        {
            const int v = std::atoi(e);
            if (v >= 1 && v <= 4)
                c->replay_big_ctas_per_sm = static_cast<uint32_t>(v);
        }
    }
    *ctx_out = c;
    return 0;
}

void lidar_b200_destroy(lidar_b200_ctx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    if (c->stream_big)
        cudaStreamSynchronize(c->stream_big);
    if (c->stream_huge)
        cudaStreamSynchronize(c->stream_huge);
    if (c->stream_small)
        cudaStreamSynchronize(c->stream_small);
    if (c->stream)
        cudaStreamSynchronize(c->stream);
    cudaFreeHost(c->h_pts.p);
    for (auto &h : c->h_u32)
        cudaFreeHost(h.p);
    cudaFreeHost(c->h_meta.p);
    cudaFreeHost(c->h_err.p);
    void *dev[] = {c->d_pts.p,      c->d_spts.p,   c->d_obs.p,       c->d_nodes.p,  c->d_cpts.p,       c->d_key_a.p,
                   c->d_key_b.p,    c->d_val_a.p,  c->d_val_b.p,     c->d_labels.p, c->d_gidx.p,       c->d_oidx.p,
                   c->d_slot_of.p,  c->d_pos_of.p, c->d_parent.p,    c->d_root.p,   c->d_rank.p,       c->d_gepos.p,
                   c->d_lepos.p,    c->d_state.p,  c->d_seed_of.p,   c->d_member_pos.p, c->d_queue.p,  c->d_seed_label.p, c->d_comp_size.p, c->d_pslot.p, c->d_rpts.p, c->d_tlive.p,
                   c->d_clabels.p,  c->d_spill.p,  c->d_pkey.p,  c->d_flags.p,     c->d_seed_valid.p, c->d_tkeys.p,  c->d_tcount.p,
                   c->d_cells.p,    c->d_nbr.p, c->d_cinfo.p, c->d_biglist.p, c->d_job_stats.p, c->d_meta.p,   c->d_err.p,       c->d_planes.p, c->d_status.p,     c->d_hist.p, c->d_goff.p, c->d_hoff.p, c->d_hne.p, c->d_hnv.p, c->d_herr.p, c->d_rgb.p, c->d_color.p, c->d_marker.p, c->d_ipts.p, c->d_mpts.p, c->d_rankpos.p, c->d_mkey.p, c->d_nb27.p, c->d_chi.p, c->d_chi_meta.p, c->d_chi_stats.p};
    for (void *p : dev)
        if (p)
            cudaFree(p);
    for (auto &e : c->ev_stage)
        if (e)
            cudaEventDestroy(e);
    drop_graphs(c);
    if (c->ev_region_a)
        cudaEventDestroy(c->ev_region_a);
    if (c->ev_region_b)
        cudaEventDestroy(c->ev_region_b);
    if (c->ev_start)
        cudaEventDestroy(c->ev_start);
    if (c->ev_stop)
        cudaEventDestroy(c->ev_stop);
    if (c->ev_fork)
        cudaEventDestroy(c->ev_fork);
    if (c->ev_join)
        cudaEventDestroy(c->ev_join);
    if (c->ev_kd)
        cudaEventDestroy(c->ev_kd);
    if (c->ev_counts)
        cudaEventDestroy(c->ev_counts);
    if (c->ev_join2)
        cudaEventDestroy(c->ev_join2);
    if (c->ev_join3)
        cudaEventDestroy(c->ev_join3);
    if (c->stream_small)
        cudaStreamDestroy(c->stream_small);
    if (c->stream_huge)
        cudaStreamDestroy(c->stream_huge);
    if (c->stream_big)
        cudaStreamDestroy(c->stream_big);
    if (c->stream)
        cudaStreamDestroy(c->stream);
    delete c;
}

int lidar_b200_reserve(lidar_b200_ctx *c, uint32_t max_points, uint32_t max_frames)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    return reserve(c, max_points, max_frames);
}

int lidar_b200_seg_configure(lidar_b200_ctx *c, const lidar_b200_seg_cfg *cfg)
{
    if (!c || !cfg)
        return LIDAR_B200_ERR_INVALID;
    const int rc = apply_seg_cfg(c, *cfg);
    if (rc)
        return rc;
    ++c->epoch; // kernel parameters changed: captured graphs are stale
    return reserve(c, c->cap_pts, c->cap_frames); // plane/status buffers depend on the configuration
}

int lidar_b200_clu_configure(lidar_b200_ctx *c, const lidar_b200_clu_cfg *cfg)
{
    if (!c || !cfg)
        return LIDAR_B200_ERR_INVALID;
    ++c->epoch; // kernel parameters change: captured graphs are stale
    return apply_clu_cfg(c, *cfg);
}

int lidar_b200_batch_stage(lidar_b200_ctx *c, uint32_t n_frames, const void *const *points, const uint32_t *n_points,
                           uint32_t stride_bytes)
{
    if (!c || (n_frames && (!points || !n_points)))
        return LIDAR_B200_ERR_INVALID;
    c->batch_is_cluster_only = false;
    return stage(c, n_frames, points, n_points, stride_bytes);
}

int lidar_b200_batch_run(lidar_b200_ctx *c)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    LB_CUDA(c, cudaSetDevice(c->device));
    LB_CUDA(c, cudaEventRecord(c->ev_start, c->stream));
    const bool graphable = c->use_graph && c->n_frames == 1u && c->max_n != 0u && !c->profiling && !c->want_job_stats;
    int rc = 0;
    if (!graphable)
        rc = run_frames(c);
    else
    {
        if (c->graphs_epoch != c->epoch) // buffers moved or the configuration changed since the graphs were captured
        {
            drop_graphs(c);
            c->graphs_epoch = c->epoch;
        }
        lidar_b200_ctx::GraphEntry *g = nullptr;
        for (auto &e : c->graphs)
            if (e.max_n == c->max_n)
                g = &e;
        if (!g)
        {
            c->graphs.emplace_back();
            g = &c->graphs.back();
            g->max_n = c->max_n;
        }
        if (g->exec)
        {
            // the host-side bookkeeping of run_clustering, then the whole frame as one launch
            c->clu_pts = c->d_obs.p;
            c->clu_counts = c->m_no();
            c->clu_max_m = c->max_n;
            c->grouped = false;
            c->hulled = false;
            c->launches += g->launches;
            ++c->graph_launches;
            LB_CUDA(c, cudaGraphLaunch(g->exec, c->stream));
        }
        else if (!g->warm || g->bad)
        {
            // first frame of this geometry: run it launch by launch (this also makes every lazy allocation and
            // attribute call of the path happen outside a capture)
            rc = run_frames(c);
            if (c->graphs_epoch == c->epoch)
                g->warm = true;
            else
            {
                drop_graphs(c); // (g dangles from here on)
                c->graphs_epoch = c->epoch;
            }
        }
        else
        {
            const uint64_t launches0 = c->launches;
            cudaGraph_t graph = nullptr;
            bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok)
            {
                const int rc_cap = run_frames(c);
                const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
                ok = rc_cap == 0 && e == cudaSuccess && graph != nullptr && c->graphs_epoch == c->epoch;
            }
            if (ok)
                ok = cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
            if (graph)
                cudaGraphDestroy(graph);
            if (ok)
            {
                g->launches = c->launches - launches0;
                ++c->graph_launches;
                LB_CUDA(c, cudaGraphLaunch(g->exec, c->stream));
            }
            else
            {
                // capture refused (an operation the path performs is not capturable on this driver): this geometry
                // stays on the launch-by-launch path
                (void)cudaGetLastError();
                g->exec = nullptr;
                g->bad = true;
                c->launches = launches0;
                rc = run_frames(c);
            }
        }
    }
    if (rc)
        return rc;
    LB_CUDA(c, cudaEventRecord(c->ev_stop, c->stream));
    return 0;
}

int lidar_b200_sync(lidar_b200_ctx *c)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int lidar_b200_batch_fetch_async(lidar_b200_ctx *c, uint32_t *point_offset_out, uint32_t *seg_labels_out,
                                 uint32_t *ground_idx_out, uint32_t *n_ground_out, uint32_t *obstacle_idx_out,
                                 uint32_t *n_obstacle_out, int32_t *cluster_labels_out, uint32_t *n_clusters_out)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    if (c->fetch.pending)
    {
        const int rc = finish_fetch(c);
        if (rc)
            return rc;
    }
    LB_CUDA(c, cudaSetDevice(c->device));
    const size_t F = c->cap_frames;
    const size_t bytes = static_cast<size_t>(c->total) * 4;
    cudaStream_t s = c->stream;
    lidar_b200_ctx::FetchReq &r = c->fetch;
    r.point_offset = point_offset_out;
    r.n_ground = n_ground_out;
    r.n_obstacle = n_obstacle_out;
    r.n_clusters = n_clusters_out;
    r.out[0] = seg_labels_out;
    r.out[1] = ground_idx_out;
    r.out[2] = obstacle_idx_out;
    r.out[3] = cluster_labels_out;
    // Two phases, so that only the used part of every result slot crosses the bus: the per-frame counts come
    // first (3 words per frame); the arrays follow with exact sizes once the counts are on the host — enqueued
    // by enqueue_fetch_payload() as soon as somebody looks (the pipeline polls, _wait blocks).
    if (c->n_frames)
        LB_CUDA(c, cudaMemcpyAsync(c->h_meta.p + 4 * F, c->m_ng(), 3 * F * 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaMemcpyAsync(c->h_err.p, c->d_err.p, 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaEventRecord(c->ev_counts, s));
    for (int k = 0; k < 4; ++k)
        r.direct[k] = bytes && r.out[k] && is_pinned(r.out[k]); // page-locked destination: the copy engine writes it directly
    r.payload_enqueued = false;
    if (c->fetch_mode == 0)
    {
        // one phase: the whole slot range of every array, enqueued right behind the kernels
        const void *dev[4] = {c->d_labels.p, c->d_gidx.p, c->d_oidx.p, c->d_clabels.p};
        for (int k = 0; k < 4; ++k)
            if (bytes && r.out[k])
                LB_CUDA(c, cudaMemcpyAsync(r.direct[k] ? r.out[k] : static_cast<void *>(c->h_u32[k].p), dev[k], bytes,
                                           cudaMemcpyDeviceToHost, s));
        r.payload_enqueued = true;
    }
    if (c->fetch_mode == 4 && c->n_frames && bytes)
    {
        // one launch writes the used part of every result slot straight into page-locked host memory (the caller's
        // arrays when they are page-locked, else the library's staging buffers)
        void *dst[4];
        bool ok = true;
        for (int k = 0; k < 4; ++k)
        {
            dst[k] = r.out[k] ? device_view_of_pinned(r.direct[k] ? r.out[k] : static_cast<void *>(c->h_u32[k].p)) : nullptr;
            ok = ok && (!r.out[k] || dst[k]);
        }
        if (!ok)
            return fail(c, LIDAR_B200_ERR_CUDA, "fetch: page-locked host memory is not mapped into the device address space");
        const BatchView bv{c->m_off(), c->m_cnt(), c->n_frames};
        emit_results_kernel<<<dim3(8, c->n_frames), 256, 0, s>>>(bv, c->m_ng(), c->m_no(), c->d_labels.p, c->d_gidx.p, c->d_oidx.p,
                                                                 c->d_clabels.p, static_cast<uint32_t *>(dst[0]),
                                                                 static_cast<uint32_t *>(dst[1]), static_cast<uint32_t *>(dst[2]),
                                                                 static_cast<int32_t *>(dst[3]));
        ++c->launches;
        LB_CUDA(c, cudaGetLastError());
        r.payload_enqueued = true;
    }
    r.with_worker = r.worker_done = false;
    r.worker_rc = 0;
    r.pending = true;
    return 0;
}

int lidar_b200_batch_wait(lidar_b200_ctx *c)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    if (!c->fetch.pending)
        return 0;
    return finish_fetch(c);
}

int lidar_b200_batch_fetch(lidar_b200_ctx *c, uint32_t *point_offset_out, uint32_t *seg_labels_out,
                           uint32_t *ground_idx_out, uint32_t *n_ground_out, uint32_t *obstacle_idx_out,
                           uint32_t *n_obstacle_out, int32_t *cluster_labels_out, uint32_t *n_clusters_out)
{
    const int rc = lidar_b200_batch_fetch_async(c, point_offset_out, seg_labels_out, ground_idx_out, n_ground_out,
                                                obstacle_idx_out, n_obstacle_out, cluster_labels_out, n_clusters_out);
    return rc ? rc : lidar_b200_batch_wait(c);
}

int lidar_b200_segment(lidar_b200_ctx *c, const void *points, uint32_t n, uint32_t stride_bytes, uint32_t *labels_inout,
                       uint32_t *ground_idx_out, uint32_t *n_ground_out, uint32_t *obstacle_idx_out,
                       uint32_t *n_obstacle_out)
{
    if (!c || !n_ground_out || !n_obstacle_out || (n && (!points || !labels_inout || !ground_idx_out || !obstacle_idx_out)))
        return LIDAR_B200_ERR_INVALID;
    *n_ground_out = 0;
    *n_obstacle_out = 0;
    if (n == 0u)
        return 0; // segmentation.cpp:319-323
    const void *frames[1] = {points};
    int rc = stage(c, 1u, frames, &n, stride_bytes);
    if (rc)
        return rc;
    c->batch_is_cluster_only = false;
    LB_CUDA(c, cudaEventRecord(c->ev_start, c->stream));
    rc = run_segmentation(c);
    if (rc)
        return rc;
    LB_CUDA(c, cudaEventRecord(c->ev_stop, c->stream));
    const size_t F = c->cap_frames;
    const size_t bytes = static_cast<size_t>(n) * 4;
    cudaStream_t s = c->stream;
    LB_CUDA(c, cudaMemcpyAsync(c->h_meta.p + 4 * F, c->m_ng(), 2 * F * 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaMemcpyAsync(c->h_u32[0].p, c->d_labels.p, bytes, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaMemcpyAsync(c->h_u32[1].p, c->d_gidx.p, bytes, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaMemcpyAsync(c->h_u32[2].p, c->d_oidx.p, bytes, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaStreamSynchronize(s));
    (void)cudaEventElapsedTime(&c->last_run_ms, c->ev_start, c->ev_stop);
    const uint32_t ng = c->h_meta.p[4 * F], no = c->h_meta.p[5 * F];
    *n_ground_out = ng;
    *n_obstacle_out = no;
    const uint32_t *hl = c->h_u32[0].p;
    for (uint32_t i = 0; i < n; ++i) // only classified points are written (segmentation.cpp:315, 334, 341)
        if (hl[i] != LIDAR_B200_SEG_UNKNOWN)
            labels_inout[i] = hl[i];
    std::memcpy(ground_idx_out, c->h_u32[1].p, static_cast<size_t>(ng) * 4);
    std::memcpy(obstacle_idx_out, c->h_u32[2].p, static_cast<size_t>(no) * 4);
    return 0;
}

int lidar_b200_cluster(lidar_b200_ctx *c, const void *points, uint32_t m, uint32_t stride_bytes, int32_t *labels_out)
{
    if (!c || (m && (!points || !labels_out)))
        return LIDAR_B200_ERR_INVALID;
    if (m == 0u)
        return 0; // clustering.cpp:50-54
    const void *frames[1] = {points};
    int rc = stage(c, 1u, frames, &m, stride_bytes);
    if (rc)
        return rc;
    c->batch_is_cluster_only = true;
    LB_CUDA(c, cudaEventRecord(c->ev_start, c->stream));
    rc = run_clustering(c, c->d_pts.p, c->m_cnt(), m);
    if (rc)
        return rc;
    LB_CUDA(c, cudaEventRecord(c->ev_stop, c->stream));
    LB_CUDA(c, cudaMemcpyAsync(c->h_u32[3].p, c->d_clabels.p, static_cast<size_t>(m) * 4, cudaMemcpyDeviceToHost, c->stream));
    LB_CUDA(c, cudaMemcpyAsync(c->h_err.p, c->d_err.p, 4, cudaMemcpyDeviceToHost, c->stream));
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    (void)cudaEventElapsedTime(&c->last_run_ms, c->ev_start, c->ev_stop);
    if (c->h_err.p[0])
        return fail(c, LIDAR_B200_ERR_INPUT, "non-finite or out-of-range point coordinates");
    std::memcpy(labels_out, c->h_u32[3].p, static_cast<size_t>(m) * 4);
    return 0;
}

int lidar_b200_batch_group_clusters(lidar_b200_ctx *c)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    if (!c->clu_pts && c->n_frames)
        return fail(c, LIDAR_B200_ERR_INVALID, "group_clusters: no clustering result on this context");
    LB_CUDA(c, cudaSetDevice(c->device));
    const uint32_t F = c->n_frames;
    if (dev_alloc(c, c->d_goff, static_cast<size_t>(c->cap_pts) + c->cap_frames + 1u))
        return LIDAR_B200_ERR_CUDA;
    cudaStream_t s = c->stream;
    c->grouped = true;
    c->hulled = false;
    if (F == 0u)
        return 0;
    LB_CUDA(c, cudaMemsetAsync(c->d_goff.p, 0, (static_cast<size_t>(c->total) + F) * 4, s));
    const uint32_t max_m = c->clu_max_m;
    if (max_m == 0u)
        return 0;
    const BatchView bv{c->m_off(), c->clu_counts, F};
    const dim3 gp(grid_x(max_m, 256u, 2048u), F);
    // every per-point scratch array of the clustering stage is free again once the labels are out
    group_keys_kernel<<<gp, 256, 0, s>>>(bv, c->d_clabels.p, c->m_nc(), c->d_key_a.p, c->d_val_a.p);
    int rl = 0;
    const int passes = radix_sort_pairs(s, c->d_key_a.p, c->d_val_a.p, c->d_key_b.p, c->d_val_b.p, bv, max_m,
                                        ceil_log2(max_m + 1u), RadixSortScratch{c->d_hist.p}, &rl);
    const uint32_t *skeys = (passes & 1) ? c->d_key_b.p : c->d_key_a.p;
    const uint32_t *svals = (passes & 1) ? c->d_val_b.p : c->d_val_a.p;
    group_emit_kernel<<<gp, 256, 0, s>>>(c->clu_pts, bv, skeys, svals, c->m_nc(), c->d_nodes.p /* grouped points */,
                                         c->d_queue.p /* grouped source indices */, c->d_goff.p);
    c->launches += 2 + rl;
    LB_CUDA(c, cudaGetLastError());
    return 0;
}

int lidar_b200_batch_fetch_clusters(lidar_b200_ctx *c, uint32_t *n_clusters_out, uint32_t *cluster_offset_out,
                                    float *cluster_points_out, uint32_t *cluster_point_idx_out)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    if (!c->grouped)
        return fail(c, LIDAR_B200_ERR_INVALID, "fetch_clusters: call lidar_b200_batch_group_clusters first");
    LB_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint32_t F = c->n_frames;
    const size_t total = c->total;
    if (F && n_clusters_out)
        LB_CUDA(c, cudaMemcpyAsync(n_clusters_out, c->m_nc(), static_cast<size_t>(F) * 4, cudaMemcpyDeviceToHost, s));
    if (F && cluster_offset_out)
        LB_CUDA(c, cudaMemcpyAsync(cluster_offset_out, c->d_goff.p, (total + F) * 4, cudaMemcpyDeviceToHost, s));
    if (total && cluster_points_out)
        LB_CUDA(c, cudaMemcpyAsync(cluster_points_out, c->d_nodes.p, total * sizeof(float4), cudaMemcpyDeviceToHost, s));
    if (total && cluster_point_idx_out)
        LB_CUDA(c, cudaMemcpyAsync(cluster_point_idx_out, c->d_queue.p, total * 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaStreamSynchronize(s));
    return 0;
}

int lidar_b200_batch_hull_outlines(lidar_b200_ctx *c, uint32_t mode)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    if (mode != LIDAR_B200_HULL_CONVEX && mode != LIDAR_B200_HULL_CONCAVE_SMALL && mode != LIDAR_B200_HULL_CONCAVE)
        return fail(c, LIDAR_B200_ERR_INVALID, "hull_outlines: unknown mode");
    if (!c->grouped)
        return fail(c, LIDAR_B200_ERR_INVALID, "hull_outlines: call lidar_b200_batch_group_clusters first");
    LB_CUDA(c, cudaSetDevice(c->device));
    const uint32_t F = c->n_frames;
    if (dev_alloc(c, c->d_hoff, static_cast<size_t>(c->cap_pts) + c->cap_frames + 1u) ||
        dev_alloc(c, c->d_hne, static_cast<size_t>(c->cap_pts) + c->cap_frames + 1u) ||
        dev_alloc(c, c->d_hnv, 4 * static_cast<size_t>(c->cap_frames) + 2) || dev_alloc(c, c->d_herr, 4))
        return LIDAR_B200_ERR_CUDA;
    if (mode == LIDAR_B200_HULL_CONCAVE && F && c->clu_max_m &&
        (dev_alloc(c, c->d_chi, static_cast<size_t>(kChiBytesPerPoint) * c->total + 256u) || dev_alloc(c, c->d_chi_meta, 2u * kChiBuckets + 1u)))
        return LIDAR_B200_ERR_CUDA;
    cudaStream_t s = c->stream;
    c->hulled = true;
    c->hull_mode = mode;
    if (F == 0u)
        return 0;
    LB_CUDA(c, cudaMemsetAsync(c->d_hoff.p, 0, (static_cast<size_t>(c->total) + F) * 4, s));
    LB_CUDA(c, cudaMemsetAsync(c->d_hnv.p, 0, (4 * static_cast<size_t>(c->cap_frames) + 2) * 4, s)); // vertices, tasks, prefix, cursor, outlines
    LB_CUDA(c, cudaMemsetAsync(c->d_herr.p, 0, 4, s));
    if (c->clu_max_m == 0u)
        return 0;
    const size_t smem = sizeof(HullWarpSmem) * kHullWarps;
    if (!c->hull_attr_done) // opt-in to > 48 KB of dynamic shared memory (a per-device attribute: once per context)
    {
        LB_CUDA(c, cudaFuncSetAttribute(hull_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        c->hull_attr_done = true;
    }
    const BatchView bv{c->m_off(), c->clu_counts, F};
    // scratch: the per-point arrays of the clustering stage and of the group sort are free by now
    const HullView hv{c->d_nodes.p, c->d_goff.p, c->d_queue.p, c->d_key_a.p, c->d_hoff.p, c->d_key_b.p, c->d_val_a.p,
                      c->d_val_b.p, reinterpret_cast<float2 *>(c->d_pkey.p), reinterpret_cast<unsigned long long *>(c->d_cpts.p),
                      c->d_slot_of.p, reinterpret_cast<uint32_t *>(c->d_rpts.p), c->d_herr.p};
    // d_hnv: [vertices per frame | tasks per frame | task prefix (F + 1) | cursor | non-empty outlines per frame]
    uint32_t *n_tasks = c->d_hnv.p + c->cap_frames, *task_base = c->d_hnv.p + 2 * static_cast<size_t>(c->cap_frames),
             *cursor = c->d_hnv.p + 3 * static_cast<size_t>(c->cap_frames) + 1;
    hull_tasks_kernel<<<F, 256, 0, s>>>(bv, c->m_nc(), hv, mode, c->d_lepos.p, c->d_state.p, n_tasks);
    hull_task_base_kernel<<<1, 256, 0, s>>>(n_tasks, F, task_base);
    hull_sort_kernel<<<c->sm_count * 4u, 32 * kHullWarps, smem, s>>>(bv, hv, c->d_lepos.p, c->d_state.p, task_base, cursor);
    hull_scan_chain_kernel<<<c->sm_count * 16u, 128, 0, s>>>(bv, hv, c->d_lepos.p, c->d_state.p, task_base);
    uint32_t nl = 6u;
    if (mode == LIDAR_B200_HULL_CONVEX && c->clu_max_m > kHullMonotoneMax)
    {
        hull_chan_merge_kernel<<<dim3(8, F), 256, 0, s>>>(bv, c->m_nc(), hv);
        ++nl;
    }
    if (mode == LIDAR_B200_HULL_CONCAVE)
    {
        // the chi-shape of every cluster from 20 points on: one warp per cluster, largest clusters first
        uint32_t *counts = c->d_chi_meta.p, *fill = counts + kChiBuckets, *chi_cursor = counts + 2u * kChiBuckets;
        uint32_t *task_f = c->d_parent.p, *task_k = c->d_comp_size.p; // (the union-find arrays are spent)
        LB_CUDA(c, cudaMemsetAsync(counts, 0, (2u * kChiBuckets + 1u) * 4, s));
        const ChiView cv{c->d_nodes.p, c->d_goff.p, c->d_key_a.p, c->d_hoff.p, c->d_chi.p, c->d_herr.p};
        chi_bucket_kernel<<<F, 256, 0, s>>>(bv, c->m_nc(), c->d_goff.p, counts);
        chi_place_kernel<<<F, 256, 0, s>>>(bv, c->m_nc(), c->d_goff.p, counts, fill, task_f, task_k);
        if (c->chi_stats)
        {
            if (dev_alloc(c, c->d_chi_stats, static_cast<size_t>(kChiStatTasks) * 8u))
                return LIDAR_B200_ERR_CUDA;
            LB_CUDA(c, cudaMemsetAsync(c->d_chi_stats.p, 0, static_cast<size_t>(kChiStatTasks) * 64u, s));
        }
        chi_outline_kernel<<<c->sm_count * c->chi_ctas_per_sm, 32 * kChiWarps, 0, s>>>(bv, cv, counts, task_f, task_k, chi_cursor,
                                                                        c->chi_stats ? c->d_chi_stats.p : nullptr);
        nl += 3u;
    }
    hull_scan_kernel<<<F, 256, 0, s>>>(bv, c->m_nc(), c->d_hoff.p, c->d_hne.p, c->d_hnv.p,
                                       c->d_hnv.p + 3 * static_cast<size_t>(c->cap_frames) + 2, c->m_cnt(), c->d_herr.p);
    hull_emit_kernel<<<dim3(8, F), 256, 0, s>>>(bv, c->m_nc(), hv, reinterpret_cast<float2 *>(c->d_spill.p), c->d_gepos.p,
                                                mode == LIDAR_B200_HULL_CONCAVE ? kHullConcaveMin : 0xFFFFFFFFu);
    c->launches += nl;
    LB_CUDA(c, cudaGetLastError());
    return 0;
}

int lidar_b200_batch_fetch_hulls(lidar_b200_ctx *c, uint32_t *n_vertices_out, uint32_t *hull_offset_out, float *hull_xy_out,
                                 uint32_t *hull_point_idx_out)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    if (!c->hulled)
        return fail(c, LIDAR_B200_ERR_INVALID, "fetch_hulls: call lidar_b200_batch_hull_outlines first");
    LB_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint32_t F = c->n_frames;
    if (F == 0u)
        return 0;
    // counts first: only the used part of every frame slot crosses the bus
    std::vector<uint32_t> nv(F), nc(F);
    uint32_t herr = 0u;
    LB_CUDA(c, cudaMemcpyAsync(nv.data(), c->d_hnv.p, static_cast<size_t>(F) * 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaMemcpyAsync(nc.data(), c->m_nc(), static_cast<size_t>(F) * 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaMemcpyAsync(&herr, c->d_herr.p, 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaStreamSynchronize(s));
    if (herr & kHullErrSubset)
        return fail(c, LIDAR_B200_ERR_UNSUPPORTED, "hull_outlines: a cluster above ~1.04 M points exceeds the CHAN subset buffers");
    const float2 *hxy = reinterpret_cast<const float2 *>(c->d_spill.p);
    for (uint32_t f = 0; f < F; ++f)
    {
        const size_t o = c->off[f];
        if (n_vertices_out)
            n_vertices_out[f] = nv[f];
        if (hull_offset_out)
            LB_CUDA(c, cudaMemcpyAsync(hull_offset_out + o + f, c->d_hoff.p + o + f, (static_cast<size_t>(nc[f]) + 1u) * 4,
                                       cudaMemcpyDeviceToHost, s));
        if (nv[f] && hull_xy_out)
            LB_CUDA(c, cudaMemcpyAsync(hull_xy_out + 2u * o, hxy + o, static_cast<size_t>(nv[f]) * sizeof(float2),
                                       cudaMemcpyDeviceToHost, s));
        if (nv[f] && hull_point_idx_out)
            LB_CUDA(c, cudaMemcpyAsync(hull_point_idx_out + o, c->d_gepos.p + o, static_cast<size_t>(nv[f]) * 4,
                                       cudaMemcpyDeviceToHost, s));
    }
    LB_CUDA(c, cudaStreamSynchronize(s));
    // The outputs are complete either way: a cluster whose Jarvis march does not close (duplicate hull vertices in
    // different CHAN subsets; the reference never returns from such a cluster) or whose hull would outgrow it has
    // 0 vertices, every other cluster is exact. The status tells the caller that it happened.
    if (herr & (kHullErrCollinear | kHullErrDegenerate))
        return fail(c, LIDAR_B200_ERR_INPUT, herr & kHullErrCollinear
                                                 ? "hull_outlines: a cluster of 20 or more points that are all collinear in (x, y) - the reference throws "
                                                   "\"not triangulation\" (delaunator.cpp:299) - was given 0 vertices; all other outlines are valid"
                                                 : "hull_outlines: a cluster of 20 or more points that all coincide in (x, y) (the reference reads out of "
                                                   "bounds) was given 0 vertices; all other outlines are valid");
    if (herr & kHullErrEnvelope)
        return fail(c, LIDAR_B200_ERR_INPUT, "hull_outlines: a cluster has two different x or y values less than FLT_EPSILON apart; the reference's "
                                             "comparators are not a strict weak order there (convex_hull.hpp:51-73) and its outline is unspecified - "
                                             "the outlines delivered use exact comparisons");
    if (herr & kHullErrSlot)
        return fail(c, LIDAR_B200_ERR_CAPACITY, "hull_outlines: the closed outlines of a frame outgrow its slot (one vertex per staged point); "
                                                "the outlines that did not fit were given 0 vertices");
    if (herr)
        return fail(c, LIDAR_B200_ERR_INPUT, "hull_outlines: a cluster on which the reference's CHAN hull does not terminate "
                                             "(Jarvis march that never closes) was given 0 vertices; all other outlines are valid");
    return 0;
}

int lidar_b200_batch_fetch_colorized(lidar_b200_ctx *c, const uint32_t *cluster_rgb, uint64_t n_rgb, float *colorized_out)
{
    if (!c || !colorized_out || (!cluster_rgb && n_rgb))
        return LIDAR_B200_ERR_INVALID;
    if (!c->grouped)
        return fail(c, LIDAR_B200_ERR_INVALID, "fetch_colorized: call lidar_b200_batch_group_clusters first");
    LB_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint32_t F = c->n_frames;
    if (F == 0u || c->clu_max_m == 0u)
        return n_rgb ? fail(c, LIDAR_B200_ERR_INVALID, "fetch_colorized: colours given but there is no cluster") : 0;
    std::vector<uint32_t> nc(F), rgb_off(F + 1u, 0u);
    LB_CUDA(c, cudaMemcpyAsync(nc.data(), c->m_nc(), static_cast<size_t>(F) * 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaStreamSynchronize(s));
    for (uint32_t f = 0; f < F; ++f)
        rgb_off[f + 1u] = rgb_off[f] + nc[f];
    if (n_rgb != rgb_off[F])
        return fail(c, LIDAR_B200_ERR_INVALID, "fetch_colorized: n_rgb must be the number of clusters of the batch");
    if (dev_alloc(c, c->d_rgb, static_cast<size_t>(c->cap_pts) + c->cap_frames + 1u)) // K <= points per frame
        return LIDAR_B200_ERR_CUDA;
    if (n_rgb == 0u)
        return 0;
    // colour words, then the per-frame prefix behind them
    LB_CUDA(c, cudaMemcpyAsync(c->d_rgb.p, cluster_rgb, n_rgb * 4, cudaMemcpyHostToDevice, s));
    LB_CUDA(c, cudaMemcpyAsync(c->d_rgb.p + c->cap_pts, rgb_off.data(), static_cast<size_t>(F) * 4, cudaMemcpyHostToDevice, s));
    const BatchView bv{c->m_off(), c->clu_counts, F};
    // 32-byte records per grouped point, allocated on first use
    if (dev_alloc(c, c->d_color, 2 * static_cast<size_t>(c->cap_pts)))
        return LIDAR_B200_ERR_CUDA;
    colorize_kernel<<<dim3(8, F), 256, 0, s>>>(bv, c->m_nc(), c->d_nodes.p, c->d_goff.p, c->d_rgb.p, c->d_rgb.p + c->cap_pts,
                                               c->d_color.p);
    ++c->launches;
    LB_CUDA(c, cudaGetLastError());
    // valid points per frame = offset[K]: one more small read, then only the used part of every slot is copied
    std::vector<uint32_t> n_valid(F, 0u);
    for (uint32_t f = 0; f < F; ++f)
        LB_CUDA(c, cudaMemcpyAsync(&n_valid[f], c->d_goff.p + c->off[f] + f + nc[f], 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaStreamSynchronize(s));
    for (uint32_t f = 0; f < F; ++f)
        if (n_valid[f])
            LB_CUDA(c, cudaMemcpyAsync(colorized_out + 8u * static_cast<size_t>(c->off[f]), c->d_color.p + 2u * static_cast<size_t>(c->off[f]),
                                       static_cast<size_t>(n_valid[f]) * 32u, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaStreamSynchronize(s));
    return 0;
}

int lidar_b200_batch_fetch_marker_points(lidar_b200_ctx *c, uint32_t *n_markers_out, uint32_t *marker_offset_out,
                                         double *marker_points_out)
{
    if (!c || !marker_points_out)
        return LIDAR_B200_ERR_INVALID;
    if (!c->hulled)
        return fail(c, LIDAR_B200_ERR_INVALID, "fetch_marker_points: call lidar_b200_batch_hull_outlines first");
    LB_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const uint32_t F = c->n_frames;
    if (F == 0u)
        return 0;
    std::vector<uint32_t> nv(F), no(F), nc(F);
    LB_CUDA(c, cudaMemcpyAsync(nv.data(), c->d_hnv.p, static_cast<size_t>(F) * 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaMemcpyAsync(no.data(), c->d_hnv.p + 3 * static_cast<size_t>(c->cap_frames) + 2, static_cast<size_t>(F) * 4,
                               cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaMemcpyAsync(nc.data(), c->m_nc(), static_cast<size_t>(F) * 4, cudaMemcpyDeviceToHost, s));
    LB_CUDA(c, cudaStreamSynchronize(s));
    if (c->clu_max_m)
    {
        if (dev_alloc(c, c->d_marker, 6 * static_cast<size_t>(c->cap_pts))) // 3 doubles per point, two slots per grouped point
            return LIDAR_B200_ERR_CUDA;
        const BatchView bv{c->m_off(), c->clu_counts, F};
        marker_points_kernel<<<dim3(8, F), 256, 0, s>>>(bv, c->m_nc(), c->d_hoff.p, c->d_hne.p,
                                                        reinterpret_cast<const float2 *>(c->d_spill.p), c->d_marker.p);
        ++c->launches;
        LB_CUDA(c, cudaGetLastError());
    }
    for (uint32_t f = 0; f < F; ++f)
    {
        const size_t o = c->off[f];
        if (n_markers_out)
            n_markers_out[f] = no[f];
        // marker k of the frame owns points [hoff[k] + hne[k], hoff[k+1] + hne[k+1]): the host adds the two CSR arrays
        if (marker_offset_out)
            LB_CUDA(c, cudaMemcpyAsync(marker_offset_out + o + f, c->d_hne.p + o + f, (static_cast<size_t>(nc[f]) + 1u) * 4,
                                       cudaMemcpyDeviceToHost, s));
        if (nv[f])
            LB_CUDA(c, cudaMemcpyAsync(marker_points_out + 6u * o, c->d_marker.p + 6u * o,
                                       (static_cast<size_t>(nv[f]) + no[f]) * 3u * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    LB_CUDA(c, cudaStreamSynchronize(s));
    return 0;
}

int lidar_b200_last_planes(lidar_b200_ctx *c, float *planes_out, int32_t *status_out)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    const size_t per = static_cast<size_t>(c->seg.partitions) * c->seg.iterations * 4;
    if (planes_out && c->n_frames)
        LB_CUDA(c, cudaMemcpy(planes_out, c->d_planes.p, c->n_frames * per * sizeof(float), cudaMemcpyDeviceToHost));
    if (status_out && c->n_frames)
        LB_CUDA(c, cudaMemcpy(status_out, c->d_status.p, static_cast<size_t>(c->n_frames) * c->seg.partitions * 4,
                              cudaMemcpyDeviceToHost));
    return 0;
}

int lidar_b200_last_kd_rank(lidar_b200_ctx *c, uint32_t frame, uint32_t *rank_out, uint32_t capacity)
{
    if (!c || frame >= c->n_frames || !rank_out)
        return LIDAR_B200_ERR_INVALID;
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    uint32_t m = 0;
    LB_CUDA(c, cudaMemcpy(&m, (c->batch_is_cluster_only ? c->m_cnt() : c->m_no()) + frame, 4, cudaMemcpyDeviceToHost));
    if (m > capacity)
        return fail(c, LIDAR_B200_ERR_CAPACITY, "rank_out too small");
    if (m)
        LB_CUDA(c, cudaMemcpy(rank_out, c->d_rank.p + c->off[frame], static_cast<size_t>(m) * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int lidar_b200_last_cc_root(lidar_b200_ctx *c, uint32_t frame, uint32_t *root_out, uint32_t capacity)
{
    if (!c || frame >= c->n_frames || !root_out)
        return LIDAR_B200_ERR_INVALID;
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    uint32_t m = 0;
    LB_CUDA(c, cudaMemcpy(&m, (c->batch_is_cluster_only ? c->m_cnt() : c->m_no()) + frame, 4, cudaMemcpyDeviceToHost));
    if (m > capacity)
        return fail(c, LIDAR_B200_ERR_CAPACITY, "root_out too small");
    if (m)
        LB_CUDA(c, cudaMemcpy(root_out, c->d_root.p + c->off[frame], static_cast<size_t>(m) * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int lidar_b200_last_chi_stats(lidar_b200_ctx *c, uint64_t *stats_out, uint32_t capacity_tasks, uint32_t *n_tasks_out)
{
    if (!c || !n_tasks_out)
        return LIDAR_B200_ERR_INVALID;
    *n_tasks_out = 0;
    if (!c->chi_stats || !c->d_chi_stats.p || !c->d_chi_meta.p)
        return fail(c, LIDAR_B200_ERR_UNSUPPORTED, "set LIDAR_B200_CHI_STATS=1 before creating the context and run mode LIDAR_B200_HULL_CONCAVE");
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    uint32_t nb[kChiBuckets];
    LB_CUDA(c, cudaMemcpy(nb, c->d_chi_meta.p, sizeof(nb), cudaMemcpyDeviceToHost));
    uint32_t n = 0;
    for (uint32_t b = 0; b < kChiBuckets; ++b)
        n += nb[b];
    *n_tasks_out = n;
    n = n < capacity_tasks ? n : capacity_tasks;
    n = n < kChiStatTasks ? n : kChiStatTasks;
    if (n && stats_out)
        LB_CUDA(c, cudaMemcpy(stats_out, c->d_chi_stats.p, static_cast<size_t>(n) * 64, cudaMemcpyDeviceToHost));
    return 0;
}

int lidar_b200_last_replay_stats(lidar_b200_ctx *c, uint32_t *stats_out, uint32_t capacity_jobs, uint32_t *n_jobs_out)
{
    if (!c || !n_jobs_out)
        return LIDAR_B200_ERR_INVALID;
    *n_jobs_out = 0;
    if (!c->d_job_stats.p)
        return fail(c, LIDAR_B200_ERR_UNSUPPORTED, "set LIDAR_B200_REPLAY_STATS=1 before creating the context");
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    uint32_t nb[kBigBuckets];
    LB_CUDA(c, cudaMemcpy(nb, c->m_cursor() + 3, sizeof(nb), cudaMemcpyDeviceToHost));
    uint32_t n = 0;
    for (uint32_t b = 0; b < kBigBuckets; ++b)
        n += nb[b];
    *n_jobs_out = n;
    n = n < capacity_jobs ? n : capacity_jobs;
    if (n && stats_out)
        LB_CUDA(c, cudaMemcpy(stats_out, c->d_job_stats.p, static_cast<size_t>(n) * 32, cudaMemcpyDeviceToHost));
    return 0;
}

uint64_t lidar_b200_launch_count(const lidar_b200_ctx *c)
{
    return c ? c->launches : 0;
}

uint64_t lidar_b200_graph_launch_count(const lidar_b200_ctx *c)
{
    return c ? c->graph_launches : 0;
}

int lidar_b200_last_run_ms(lidar_b200_ctx *c, float *ms_out)
{
    if (!c || !ms_out)
        return LIDAR_B200_ERR_INVALID;
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.0f;
    if (cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop) != cudaSuccess)
    {
        (void)cudaGetLastError();
        ms = c->last_run_ms;
    }
    *ms_out = ms;
    return 0;
}

int lidar_b200_region_begin(lidar_b200_ctx *c)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    LB_CUDA(c, cudaSetDevice(c->device));
    if (!c->ev_region_a)
    {
        LB_CUDA(c, cudaEventCreate(&c->ev_region_a));
        LB_CUDA(c, cudaEventCreate(&c->ev_region_b));
    }
    LB_CUDA(c, cudaEventRecord(c->ev_region_a, c->stream));
    return 0;
}

int lidar_b200_region_end_ms(lidar_b200_ctx *c, float *ms_out)
{
    if (!c || !ms_out)
        return LIDAR_B200_ERR_INVALID;
    if (!c->ev_region_a)
        return fail(c, LIDAR_B200_ERR_INVALID, "region_end_ms: call lidar_b200_region_begin first");
    LB_CUDA(c, cudaSetDevice(c->device));
    // every batch_run joins its side streams back into the main stream, so an event on it closes the region
    LB_CUDA(c, cudaEventRecord(c->ev_region_b, c->stream));
    LB_CUDA(c, cudaEventSynchronize(c->ev_region_b));
    LB_CUDA(c, cudaEventElapsedTime(ms_out, c->ev_region_a, c->ev_region_b));
    return 0;
}

int lidar_b200_set_profiling(lidar_b200_ctx *c, int enabled)
{
    if (!c)
        return LIDAR_B200_ERR_INVALID;
    c->profiling = enabled != 0;
    c->n_stage_marks = 0;
    return 0;
}

int lidar_b200_last_stage_ms(lidar_b200_ctx *c, float *ms_out, uint32_t capacity)
{
    if (!c || !ms_out || capacity < static_cast<uint32_t>(lidar_b200_ctx::kStages - 1))
        return LIDAR_B200_ERR_INVALID;
    LB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i + 1 < lidar_b200_ctx::kStages; ++i)
    {
        ms_out[i] = 0.0f;
        if (c->profiling && i + 1 < c->n_stage_marks &&
            cudaEventElapsedTime(&ms_out[i], c->ev_stage[i], c->ev_stage[i + 1]) != cudaSuccess)
        {
            (void)cudaGetLastError();
            ms_out[i] = 0.0f;
        }
    }
    return 0;
}

const char *lidar_b200_last_error(const lidar_b200_ctx *c)
{
    return c ? c->err.c_str() : "null context";
}

const char *lidar_b200_version(void)
{
    return "lidar_b200 0.1 (sm_100a)";
}

int lidar_b200_pcd_read(const char *path, void *points_out, uint64_t capacity_points, uint32_t stride_bytes,
                        uint64_t *n_points_out, char *error_out, uint32_t error_capacity)
{
    const std::string err = pcd_read(path, points_out, capacity_points, stride_bytes, n_points_out);
    if (error_out && error_capacity)
    {
        std::strncpy(error_out, err.c_str(), error_capacity - 1u);
        error_out[error_capacity - 1u] = '\0';
    }
    return err.empty() ? LIDAR_B200_OK : LIDAR_B200_ERR_INVALID;
}

int lidar_b200_host_alloc(void **ptr_out, uint64_t bytes)
{
    if (!ptr_out)
        return LIDAR_B200_ERR_INVALID;
    *ptr_out = nullptr;
    if (cudaHostAlloc(ptr_out, bytes ? bytes : 1u, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess)
    {
        (void)cudaGetLastError();
        *ptr_out = nullptr;
        return LIDAR_B200_ERR_CUDA;
    }
    return 0;
}

void lidar_b200_host_free(void *ptr)
{
    if (ptr && cudaFreeHost(ptr) != cudaSuccess)
        (void)cudaGetLastError();
}

// ---- frame pipeline: `depth` contexts used round-robin, so that the upload of chunk k+1, the
// kernels of chunk k and the download of chunk k-1 overlap (three engines, independent streams)
// The pipeline's second-phase worker: waits for the counts of a submitted chunk and enqueues its exact-size result
// copies on the chunk's own stream, so the submitting thread never blocks on a chunk that is still running and the
// copies start the moment the chunk's kernels end.
struct lidar_b200_pipe
{
    std::vector<lidar_b200_ctx *> slots;
    uint32_t next{0};
    std::string err;
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<lidar_b200_ctx *> work;
    bool stop{false};
    // Result fetch mode of the pipe: LIDAR_B200_FETCH_MODE pins it, else lidar_b200_pipe_set_host_sharing decides.
    bool mode_pinned{false};
    int mode{0};
};

namespace
{
void pipe_worker(lidar_b200_pipe *p)
{
    std::unique_lock<std::mutex> lk(p->mu);
    while (true)
    {
        p->cv_work.wait(lk, [p] { return p->stop || !p->work.empty(); });
        if (p->work.empty())
            return; // stop requested and nothing left
        lidar_b200_ctx *c = p->work.front();
        p->work.pop_front();
        lk.unlock();
        const int rc = enqueue_fetch_payload(c, true);
        lk.lock();
        c->fetch.worker_rc = rc;
        c->fetch.worker_done = true;
        p->cv_done.notify_all();
    }
}

// the submitting thread: before touching a slot again, let the worker finish its part of the slot's last chunk
int pipe_wait_worker(lidar_b200_pipe *p, lidar_b200_ctx *c)
{
    std::unique_lock<std::mutex> lk(p->mu);
    if (!c->fetch.pending || !c->fetch.with_worker)
        return 0;
    p->cv_done.wait(lk, [c] { return c->fetch.worker_done; });
    return c->fetch.worker_rc;
}
} // namespace

int lidar_b200_pipe_create(int device, uint32_t depth, uint32_t max_points, uint32_t max_frames, lidar_b200_pipe **pipe_out)
{
    if (!pipe_out || depth == 0u || depth > 16u)
        return LIDAR_B200_ERR_INVALID;
    *pipe_out = nullptr;
    lidar_b200_pipe *p = new lidar_b200_pipe();
    for (uint32_t i = 0; i < depth; ++i)
    {
        lidar_b200_ctx *c = nullptr;
        const int rc = lidar_b200_create(device, max_points, max_frames, &c);
        if (rc)
        {
            lidar_b200_pipe_destroy(p);
            return rc;
        }
        p->slots.push_back(c);
    }
    p->worker = std::thread(pipe_worker, p);
    p->mode_pinned = std::getenv("LIDAR_B200_FETCH_MODE") != nullptr;
    p->mode = p->slots.front()->fetch_mode;
    *pipe_out = p;
    return 0;
}

// How many GPUs of this host run a pipe at the same time (the caller knows: its local world size). The host's DMA path is
// shared: measured on the 8-GPU box (profiles/README.md, "copy-only ceiling") the copy engines of 8 GPUs move 124 GB/s in
// total, so beyond 2 GPUs the bytes decide - mode 4 (exact sizes, 39 % fewer D2H bytes, posted writes from one kernel) is
// +18 % end to end at 4 GPUs and +38 % at 8; at 1-2 GPUs the copy engines are not the limit and the slot-size copies of
// mode 0, which cost no SM time, are 10 % faster.
int lidar_b200_pipe_set_host_sharing(lidar_b200_pipe *p, uint32_t gpus_sharing_host)
{
    if (!p)
        return LIDAR_B200_ERR_INVALID;
    if (!p->mode_pinned)
        p->mode = gpus_sharing_host >= 3u ? 4 : 0;
    return 0;
}

int lidar_b200_pipe_fetch_mode(const lidar_b200_pipe *p)
{
    return p ? p->mode : -1;
}

void lidar_b200_pipe_destroy(lidar_b200_pipe *p)
{
    if (!p)
        return;
    if (p->worker.joinable())
    {
        {
            std::lock_guard<std::mutex> lk(p->mu);
            p->stop = true;
        }
        p->cv_work.notify_all();
        p->worker.join(); // drains the queue first
    }
    for (lidar_b200_ctx *c : p->slots)
        lidar_b200_destroy(c);
    delete p;
}

int lidar_b200_pipe_seg_configure(lidar_b200_pipe *p, const lidar_b200_seg_cfg *cfg)
{
    if (!p)
        return LIDAR_B200_ERR_INVALID;
    for (lidar_b200_ctx *c : p->slots)
    {
        const int rc = lidar_b200_seg_configure(c, cfg);
        if (rc)
        {
            p->err = c->err;
            return rc;
        }
    }
    return 0;
}

int lidar_b200_pipe_clu_configure(lidar_b200_pipe *p, const lidar_b200_clu_cfg *cfg)
{
    if (!p)
        return LIDAR_B200_ERR_INVALID;
    for (lidar_b200_ctx *c : p->slots)
    {
        const int rc = lidar_b200_clu_configure(c, cfg);
        if (rc)
        {
            p->err = c->err;
            return rc;
        }
    }
    return 0;
}

int lidar_b200_pipe_submit(lidar_b200_pipe *p, uint32_t n_frames, const void *const *points, const uint32_t *n_points,
                           uint32_t stride_bytes, uint32_t *point_offset_out, uint32_t *seg_labels_out,
                           uint32_t *ground_idx_out, uint32_t *n_ground_out, uint32_t *obstacle_idx_out,
                           uint32_t *n_obstacle_out, int32_t *cluster_labels_out, uint32_t *n_clusters_out)
{
    if (!p || p->slots.empty())
        return LIDAR_B200_ERR_INVALID;
    lidar_b200_ctx *c = p->slots[p->next];
    p->next = (p->next + 1u) % static_cast<uint32_t>(p->slots.size());
    // staging completes the chunk this slot still owes (its results are in caller memory afterwards)
    int rc = pipe_wait_worker(p, c);
    if (!rc)
        rc = lidar_b200_batch_stage(c, n_frames, points, n_points, stride_bytes);
    if (!rc)
        rc = lidar_b200_batch_run(c);
    // (mode 4 needs mapped page-locked destinations; a chunk with pageable result arrays is fetched with the copies)
    c->fetch_mode = p->mode;
    if (p->mode == 4 && !p->mode_pinned &&
        !(device_view_of_pinned(seg_labels_out) && device_view_of_pinned(obstacle_idx_out) && device_view_of_pinned(cluster_labels_out) &&
          (!ground_idx_out || device_view_of_pinned(ground_idx_out))))
        c->fetch_mode = 0;
    if (!rc)
        rc = lidar_b200_batch_fetch_async(c, point_offset_out, seg_labels_out, ground_idx_out, n_ground_out,
                                          obstacle_idx_out, n_obstacle_out, cluster_labels_out, n_clusters_out);
    if (!rc && !c->fetch.payload_enqueued)
    {
        // hand the second phase (exact-size result copies once the counts are on the host) to the worker
        std::lock_guard<std::mutex> lk(p->mu);
        c->fetch.with_worker = true;
        p->work.push_back(c);
        p->cv_work.notify_one();
    }
    if (rc)
        p->err = c->err;
    return rc;
}

int lidar_b200_pipe_drain(lidar_b200_pipe *p)
{
    if (!p)
        return LIDAR_B200_ERR_INVALID;
    int first = 0;
    const uint32_t k = static_cast<uint32_t>(p->slots.size());
    for (uint32_t i = 0; i < k; ++i) // oldest chunk first
    {
        lidar_b200_ctx *c = p->slots[(p->next + i) % k];
        int rc = pipe_wait_worker(p, c);
        const int rc2 = lidar_b200_batch_wait(c);
        rc = rc ? rc : rc2;
        if (rc && !first)
        {
            first = rc;
            p->err = c->err;
        }
    }
    return first;
}

uint64_t lidar_b200_pipe_launch_count(const lidar_b200_pipe *p)
{
    uint64_t n = 0;
    if (p)
        for (const lidar_b200_ctx *c : p->slots)
            n += c->launches;
    return n;
}

const char *lidar_b200_pipe_last_error(const lidar_b200_pipe *p)
{
    return p ? p->err.c_str() : "null pipe";
}

} // extern "C"
