"""lidar_b200 — B200-native ground segmentation + Fast Euclidean Clustering.

Host-side mirror of the reference's operator interface for this path
(`lidar_processing::Segmenter`, `lidar_processing::Clusterer`; reference src/segmentation.hpp:58-70,
src/clustering.hpp:50-64) on top of the C ABI in include/lidar_b200.h. The product is the CUDA
library `liblidar_b200.so` built from csrc/; this module only marshals numpy / torch buffers into
it. There is no CPU fallback: constructing any class without the built library or without a CUDA
device raises.

The directory name contains a hyphen, so import it through `__graft_entry__.load_package()` (or
importlib) under the module name `lidar_processing_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import weakref
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
LIB_PATH = HERE / "liblidar_b200.so"
CSRC = HERE / "csrc"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # the reference build has no FMA contraction; parity depends on it
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread",
]

UNKNOWN, GROUND, OBSTACLE = 0, 1, 2
UNDEFINED = np.iinfo(np.int32).min
INVALID = -1

_lib = None


class LidarB200Error(RuntimeError):
    pass


class SegmentationConfiguration(C.Structure):
    """reference src/segmentation.hpp:48-56 (same fields, same defaults)"""

    _fields_ = [
        ("sensor_height_m", C.c_float),
        ("orthogonal_distance_threshold", C.c_float),
        ("initial_seed_threshold", C.c_float),
        ("number_of_iterations", C.c_uint32),
        ("number_of_planar_partitions", C.c_uint32),
        ("number_of_lower_point_representatives", C.c_uint32),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.sensor_height_m = 1.73
        self.orthogonal_distance_threshold = 0.3
        self.initial_seed_threshold = 0.6
        self.number_of_iterations = 3
        self.number_of_planar_partitions = 2
        self.number_of_lower_point_representatives = 5000
        for k, v in kw.items():
            setattr(self, k, v)


class ClusteringConfiguration(C.Structure):
    """reference src/clustering.hpp:42-48 (same fields, same defaults)"""

    _fields_ = [
        ("distance_squared", C.c_float),
        ("cluster_quality", C.c_float),
        ("min_cluster_size", C.c_uint32),
        ("max_cluster_size", C.c_uint32),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.distance_squared = 0.18
        self.cluster_quality = 0.5
        self.min_cluster_size = 4
        self.max_cluster_size = 0xFFFFFFFF
        for k, v in kw.items():
            setattr(self, k, v)


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "lidar_b200.h"]


def build_extension(force: bool = False, verbose: bool = False) -> Path:
    """nvcc cross-compiles the CUDA library for sm_100a in-tree (no GPU needed to build)."""
    newest = max(p.stat().st_mtime for p in sources())
    if not force and LIB_PATH.exists() and LIB_PATH.stat().st_mtime >= newest:
        return LIB_PATH
    cmd = ["nvcc", *NVCC_FLAGS, "-o", str(LIB_PATH), str(CSRC / "api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise LidarB200Error(f"{LIB_PATH} is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(str(LIB_PATH))
        L.lidar_b200_last_error.restype = C.c_char_p
        L.lidar_b200_version.restype = C.c_char_p
        L.lidar_b200_launch_count.restype = C.c_uint64
        L.lidar_b200_graph_launch_count.restype = C.c_uint64
        L.lidar_b200_pipe_launch_count.restype = C.c_uint64
        L.lidar_b200_pipe_last_error.restype = C.c_char_p
        L.lidar_b200_host_free.restype = None
        _lib = L
    return _lib


EXPORTED_SYMBOLS = [
    "lidar_b200_seg_cfg_default", "lidar_b200_clu_cfg_default", "lidar_b200_create", "lidar_b200_destroy",
    "lidar_b200_reserve", "lidar_b200_seg_configure", "lidar_b200_clu_configure", "lidar_b200_segment",
    "lidar_b200_cluster", "lidar_b200_batch_stage", "lidar_b200_batch_run", "lidar_b200_batch_fetch",
    "lidar_b200_sync", "lidar_b200_last_planes", "lidar_b200_last_kd_rank", "lidar_b200_last_cc_root",
    "lidar_b200_launch_count", "lidar_b200_graph_launch_count", "lidar_b200_last_run_ms", "lidar_b200_region_begin", "lidar_b200_region_end_ms", "lidar_b200_last_error", "lidar_b200_version",
    "lidar_b200_set_profiling", "lidar_b200_last_stage_ms", "lidar_b200_batch_fetch_async", "lidar_b200_batch_wait",
    "lidar_b200_host_alloc", "lidar_b200_host_free", "lidar_b200_pipe_create", "lidar_b200_pipe_destroy",
    "lidar_b200_pipe_seg_configure", "lidar_b200_pipe_clu_configure", "lidar_b200_pipe_submit",
    "lidar_b200_pipe_drain", "lidar_b200_pipe_launch_count", "lidar_b200_pipe_last_error", "lidar_b200_pipe_fetch_mode",
    "lidar_b200_pipe_set_host_sharing",
    "lidar_b200_last_replay_stats", "lidar_b200_last_chi_stats", "lidar_b200_batch_group_clusters", "lidar_b200_batch_fetch_clusters",
    "lidar_b200_pcd_read", "lidar_b200_batch_hull_outlines", "lidar_b200_batch_fetch_hulls",
    "lidar_b200_batch_fetch_colorized", "lidar_b200_batch_fetch_marker_points",
]

DEFAULT_FETCH_MODE = 0   # the library's default LIDAR_B200_FETCH_MODE (api.cu)
ERR_INPUT = 5            # LIDAR_B200_ERR_INPUT (include/lidar_b200.h)
HULL_CONVEX = 0          # findOrderedConvexOutlines (reference src/polygon_simplification.cpp:31-79)
HULL_CONCAVE_SMALL = 1   # convex branch of findOrderedConcaveOutlines (:100-118); >= 20 points stay on the host
HULL_CONCAVE = 2         # findOrderedConcaveOutlines as a whole (:81-149): + the Delaunay-based chi-shape from 20 points on, closed


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """numpy array in page-locked host memory (lidar_b200_host_alloc): clouds and result arrays kept
    there are read / written by the copy engines directly, without a staging pass."""
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
    ptr = C.c_void_p()
    rc = lib().lidar_b200_host_alloc(C.byref(ptr), C.c_uint64(max(nbytes, 1)))
    if rc != 0 or not ptr.value:
        raise LidarB200Error(f"lidar_b200_host_alloc({nbytes}) failed (status {rc})")
    buf = (C.c_byte * max(nbytes, 1)).from_address(ptr.value)
    weakref.finalize(buf, lib().lidar_b200_host_free, C.c_void_p(ptr.value))
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


def pin_frames(frames):
    """Copies a list of (N, 4) float32 clouds into one page-locked arena; returns the list of views."""
    frames = [_as_points(f) for f in frames]
    if any(f.shape[1] != 4 for f in frames):
        raise ValueError("pin_frames expects 16-byte records (N, 4) float32")
    total = sum(f.shape[0] for f in frames)
    arena = pinned_empty((max(total, 1), 4), np.float32)
    out, pos = [], 0
    for f in frames:
        v = arena[pos:pos + f.shape[0]]
        v[...] = f
        out.append(v)
        pos += f.shape[0]
    return out


def _as_points(points) -> np.ndarray:
    a = np.asarray(points)
    if a.dtype != np.float32 or a.ndim != 2 or a.shape[1] < 3 or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError("points must be (N, >=3) float32: x, y, z first")
    return a


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def read_pcd(path, stride_bytes: int = 16) -> np.ndarray:
    """PCD v0.7 file -> (N, 4) float32 x, y, z, intensity (stride 16) or (N, 8) float32 in the 32-byte
    pcl::PointXYZI wire layout (reference src/dataloader.cpp:87-126, 139). Host code of the library."""
    n = C.c_uint64(0)
    err = C.create_string_buffer(256)
    path_b = str(path).encode()
    if lib().lidar_b200_pcd_read(path_b, None, C.c_uint64(0), C.c_uint32(stride_bytes), C.byref(n), err, C.c_uint32(256)):
        raise LidarB200Error(f"read_pcd({path}): {err.value.decode()}")
    out = np.zeros((max(int(n.value), 1), stride_bytes // 4), np.float32)
    if lib().lidar_b200_pcd_read(path_b, out.ctypes.data_as(C.c_void_p), C.c_uint64(n.value), C.c_uint32(stride_bytes),
                                 C.byref(n), err, C.c_uint32(256)):
        raise LidarB200Error(f"read_pcd({path}): {err.value.decode()}")
    return out[:int(n.value)]


class Context:
    """Owns one lidar_b200_ctx (stream, pinned staging, device arenas) on one GPU."""

    def __init__(self, device: int = 0, max_points: int = 200_000, max_frames: int = 1):
        self._h = C.c_void_p()
        rc = lib().lidar_b200_create(C.c_int(device), C.c_uint32(max_points), C.c_uint32(max_frames), C.byref(self._h))
        if rc != 0:
            raise LidarB200Error(f"lidar_b200_create failed (status {rc}): CUDA device {device} unavailable; "
                                 "there is no CPU fallback")
        self.device = device
        self.seg_cfg = SegmentationConfiguration()
        self.clu_cfg = ClusteringConfiguration()
        self._n_points = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().lidar_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = lib().lidar_b200_last_error(self._h)
            raise LidarB200Error(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")

    def reserve(self, max_points: int, max_frames: int = 1):
        self._check(lib().lidar_b200_reserve(self._h, C.c_uint32(max_points), C.c_uint32(max_frames)), "reserve")

    def seg_configure(self, cfg: SegmentationConfiguration):
        self._check(lib().lidar_b200_seg_configure(self._h, C.byref(cfg)), "seg_configure")
        self.seg_cfg = cfg

    def clu_configure(self, cfg: ClusteringConfiguration):
        self._check(lib().lidar_b200_clu_configure(self._h, C.byref(cfg)), "clu_configure")
        self.clu_cfg = cfg

    # -- single-frame calls (the reference's per-frame interface) -------------------------------
    def segment(self, points, labels_inout: np.ndarray | None = None):
        pts = _as_points(points)
        n = pts.shape[0]
        labels = np.zeros(n, np.uint32) if labels_inout is None else labels_inout
        assert labels.dtype == np.uint32 and labels.size >= n
        g = np.empty(max(n, 1), np.uint32)
        o = np.empty(max(n, 1), np.uint32)
        ng, no = C.c_uint32(0), C.c_uint32(0)
        self._check(lib().lidar_b200_segment(self._h, pts.ctypes.data_as(C.c_void_p), C.c_uint32(n),
                                             C.c_uint32(pts.shape[1] * 4), _ptr(labels, C.c_uint32),
                                             _ptr(g, C.c_uint32), C.byref(ng), _ptr(o, C.c_uint32), C.byref(no)),
                    "segment")
        return labels, g[: ng.value].copy(), o[: no.value].copy()

    def cluster(self, points) -> np.ndarray:
        pts = _as_points(points)
        m = pts.shape[0]
        labels = np.full(max(m, 1), UNDEFINED, np.int32)
        self._check(lib().lidar_b200_cluster(self._h, pts.ctypes.data_as(C.c_void_p), C.c_uint32(m),
                                             C.c_uint32(pts.shape[1] * 4), _ptr(labels, C.c_int32)), "cluster")
        return labels[:m]

    # -- batched frame pipeline -----------------------------------------------------------------
    def batch_stage(self, frames):
        frames = [_as_points(f) for f in frames]
        nf = len(frames)
        strides = {f.shape[1] * 4 for f in frames} or {16}
        if len(strides) != 1:
            raise ValueError("all frames of a batch must share one point stride")
        ptrs = (C.c_void_p * max(nf, 1))(*[f.ctypes.data for f in frames])
        counts = np.array([f.shape[0] for f in frames], np.uint32)
        self._keep = frames
        self._n_points = counts
        self._check(lib().lidar_b200_batch_stage(self._h, C.c_uint32(nf), ptrs, _ptr(counts, C.c_uint32),
                                                 C.c_uint32(strides.pop())), "batch_stage")

    def batch_run(self):
        self._check(lib().lidar_b200_batch_run(self._h), "batch_run")

    def sync(self):
        self._check(lib().lidar_b200_sync(self._h), "sync")

    def batch_fetch(self, want_ground_idx: bool = True):
        counts = self._n_points
        nf = counts.size
        total = int(((counts.astype(np.int64) + 31) & ~31).sum())
        off = np.zeros(max(nf, 1), np.uint32)
        seg = np.empty(max(total, 1), np.uint32)
        gidx = np.empty(max(total, 1), np.uint32) if want_ground_idx else None
        oidx = np.empty(max(total, 1), np.uint32)
        clab = np.empty(max(total, 1), np.int32)
        ng = np.zeros(max(nf, 1), np.uint32)
        no = np.zeros(max(nf, 1), np.uint32)
        nc = np.zeros(max(nf, 1), np.uint32)
        self._check(lib().lidar_b200_batch_fetch(self._h, _ptr(off, C.c_uint32), _ptr(seg, C.c_uint32),
                                                 _ptr(gidx, C.c_uint32) if want_ground_idx else None,
                                                 _ptr(ng, C.c_uint32), _ptr(oidx, C.c_uint32), _ptr(no, C.c_uint32),
                                                 _ptr(clab, C.c_int32), _ptr(nc, C.c_uint32)), "batch_fetch")
        out = []
        for f in range(nf):
            o, n = int(off[f]), int(counts[f])
            out.append(dict(
                seg_labels=seg[o:o + n],
                ground_idx=gidx[o:o + int(ng[f])] if want_ground_idx else None,
                obstacle_idx=oidx[o:o + int(no[f])],
                cluster_labels=clab[o:o + int(no[f])],
                n_clusters=int(nc[f]),
            ))
        return out

    def batch_clusters(self):
        """Per-cluster point compaction on the device (reference src/processor.cpp:180-200) of the last
        batch_run / cluster call. Returns, per frame, dict(offsets[K+1], points[n_valid,4], point_idx[n_valid]):
        cluster k = points[offsets[k]:offsets[k+1]] in ascending obstacle-cloud index."""
        counts = self._n_points
        nf = counts.size
        padded = ((counts.astype(np.int64) + 31) & ~31)
        total = int(padded.sum())
        off = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64) if nf else np.zeros(0, np.int64)
        nc = np.zeros(max(nf, 1), np.uint32)
        goff = np.zeros(max(total + nf, 1), np.uint32)
        gpts = np.zeros((max(total, 1), 4), np.float32)
        gidx = np.zeros(max(total, 1), np.uint32)
        self._check(lib().lidar_b200_batch_group_clusters(self._h), "batch_group_clusters")
        self._check(lib().lidar_b200_batch_fetch_clusters(self._h, _ptr(nc, C.c_uint32), _ptr(goff, C.c_uint32),
                                                          _ptr(gpts, C.c_float), _ptr(gidx, C.c_uint32)), "batch_fetch_clusters")
        out = []
        for f in range(nf):
            o, k = int(off[f]), int(nc[f])
            offsets = goff[o + f:o + f + k + 1]
            nv = int(offsets[k]) if offsets.size else 0
            out.append(dict(offsets=offsets, points=gpts[o:o + nv], point_idx=gidx[o:o + nv], n_clusters=k))
        return out

    def batch_hulls(self, mode: int = HULL_CONVEX, tolerate_open_marches: bool = False):
        """Ordered outlines per cluster on the device (reference src/polygon_simplification.cpp:31-79 / :81-149) of
        the last batch_clusters(). Returns, per frame, dict(offsets[K+1], xy[n_vertices,2], point_idx[n_vertices]):
        outline of cluster k = xy[offsets[k]:offsets[k+1]]; convex outlines are counter-clockwise and open, the concave
        ones of mode HULL_CONCAVE (clusters from 20 points on) are closed like the reference's."""
        counts = self._n_points
        nf = counts.size
        padded = ((counts.astype(np.int64) + 31) & ~31)
        total = int(padded.sum())
        off = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64) if nf else np.zeros(0, np.int64)
        nv = np.zeros(max(nf, 1), np.uint32)
        nc = np.zeros(max(nf, 1), np.uint32)
        hoff = np.zeros(max(total + nf, 1), np.uint32)
        hxy = np.zeros((max(total, 1), 2), np.float32)
        hidx = np.zeros(max(total, 1), np.uint32)
        self._check(lib().lidar_b200_batch_hull_outlines(self._h, C.c_uint32(mode)), "batch_hull_outlines")
        rc = lib().lidar_b200_batch_fetch_hulls(self._h, _ptr(nv, C.c_uint32), _ptr(hoff, C.c_uint32),
                                                _ptr(hxy, C.c_float), _ptr(hidx, C.c_uint32))
        self.last_hull_status = rc
        if not (tolerate_open_marches and rc == ERR_INPUT):  # ERR_INPUT: outputs complete, the offending clusters are empty
            self._check(rc, "batch_fetch_hulls")
        self._check(lib().lidar_b200_batch_fetch_clusters(self._h, _ptr(nc, C.c_uint32), None, None, None), "batch_fetch_clusters")
        out = []
        for f in range(nf):
            o, k, n = int(off[f]), int(nc[f]), int(nv[f])
            out.append(dict(offsets=hoff[o + f:o + f + k + 1], xy=hxy[o:o + n], point_idx=hidx[o:o + n], n_clusters=k))
        return out

    def last_chi_stats(self, capacity: int = 4096):
        """Per-cluster timings of the last batch_hulls(HULL_CONCAVE) (needs LIDAR_B200_CHI_STATS=1 at context creation):
        (tasks, 8) uint64 = points, start ns, end ns, cycles of seed / sort / sweep / erosion, triangles."""
        st = np.zeros((capacity, 8), np.uint64)
        n = C.c_uint32(0)
        self._check(lib().lidar_b200_last_chi_stats(self._h, _ptr(st, C.c_uint64), C.c_uint32(capacity), C.byref(n)), "last_chi_stats")
        return st[:min(int(n.value), capacity)], int(n.value)

    def _slots(self):
        counts = self._n_points
        padded = ((counts.astype(np.int64) + 31) & ~31)
        off = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64) if counts.size else np.zeros(0, np.int64)
        return counts.size, int(padded.sum()), off

    def batch_colorized(self, cluster_rgb, groups):
        """convertClusteredCloudToColorizedCloud on the device (reference src/conversions.cpp:32-60) for the clusters of
        the last batch_clusters() (`groups` = its result). cluster_rgb: uint32 r << 16 | g << 8 | b per cluster, frames end
        to end. Returns, per frame, the (n_valid, 32) uint8 PointXYZRGB records."""
        nf, total, off = self._slots()
        rgb = np.ascontiguousarray(cluster_rgb, np.uint32)
        out = np.zeros((max(total, 1), 8), np.float32)
        self._check(lib().lidar_b200_batch_fetch_colorized(self._h, _ptr(rgb, C.c_uint32) if rgb.size else None,
                                                           C.c_uint64(rgb.size), _ptr(out, C.c_float)), "batch_fetch_colorized")
        res = []
        for f in range(nf):
            nv = int(groups[f]["offsets"][-1]) if groups[f]["offsets"].size else 0
            res.append(out[int(off[f]):int(off[f]) + nv].view(np.uint8).reshape(nv, 32))
        return res

    def batch_marker_points(self, hulls):
        """Points of convertPointXYZTypeToMarkerArray (reference src/conversions.hpp:72-120) for the outlines of the last
        batch_hulls() (`hulls` = its result). Returns, per frame, dict(points[n,3] float64, offsets[K+1]): the marker of
        cluster k = points[offsets[k]:offsets[k+1]] (empty outline: no marker)."""
        nf, total, off = self._slots()
        nm = np.zeros(max(nf, 1), np.uint32)
        moff = np.zeros(max(total + nf, 1), np.uint32)
        pts = np.zeros((max(2 * total, 1), 3), np.float64)
        self._check(lib().lidar_b200_batch_fetch_marker_points(self._h, _ptr(nm, C.c_uint32), _ptr(moff, C.c_uint32),
                                                               _ptr(pts, C.c_double)), "batch_fetch_marker_points")
        res = []
        for f in range(nf):
            o, k = int(off[f]), hulls[f]["n_clusters"]
            offsets = hulls[f]["offsets"].astype(np.int64) + moff[o + f:o + f + k + 1].astype(np.int64) if k else np.zeros(1, np.int64)
            res.append(dict(points=pts[2 * o:2 * o + int(offsets[-1])], offsets=offsets, n_markers=int(nm[f])))
        return res

    def cluster_and_split(self, points):
        """Clusterer::cluster followed by the device-side split (single frame)."""
        pts = _as_points(points)
        labels = self.cluster(pts)
        if pts.shape[0] == 0:
            return labels, dict(offsets=np.zeros(1, np.uint32), points=np.zeros((0, 4), np.float32),
                                point_idx=np.zeros(0, np.uint32), n_clusters=0)
        self._n_points = np.array([pts.shape[0]], np.uint32)
        return labels, self.batch_clusters()[0]

    def process_batch(self, frames, want_ground_idx: bool = True):
        self.batch_stage(frames)
        self.batch_run()
        return self.batch_fetch(want_ground_idx)

    # -- diagnostics ----------------------------------------------------------------------------
    def last_planes(self, n_frames: int = 1):
        P, it = self.seg_cfg.number_of_planar_partitions, self.seg_cfg.number_of_iterations
        planes = np.zeros((n_frames, P, it, 4), np.float32)
        status = np.zeros((n_frames, P), np.int32)
        self._check(lib().lidar_b200_last_planes(self._h, _ptr(planes, C.c_float), _ptr(status, C.c_int32)),
                    "last_planes")
        return planes, status

    def last_kd_rank(self, m: int, frame: int = 0) -> np.ndarray:
        r = np.zeros(max(m, 1), np.uint32)
        self._check(lib().lidar_b200_last_kd_rank(self._h, C.c_uint32(frame), _ptr(r, C.c_uint32), C.c_uint32(m)),
                    "last_kd_rank")
        return r[:m]

    def last_cc_root(self, m: int, frame: int = 0) -> np.ndarray:
        r = np.zeros(max(m, 1), np.uint32)
        self._check(lib().lidar_b200_last_cc_root(self._h, C.c_uint32(frame), _ptr(r, C.c_uint32), C.c_uint32(m)),
                    "last_cc_root")
        return r[:m]

    STAGES = ("x_sort", "gather_fit", "compact", "voxel_grid", "union_find", "component_sort", "kd_order", "replay",
              "label_compact")

    def set_profiling(self, enabled: bool):
        self._check(lib().lidar_b200_set_profiling(self._h, C.c_int(1 if enabled else 0)), "set_profiling")

    def last_stage_ms(self) -> dict:
        ms = (C.c_float * 9)()
        self._check(lib().lidar_b200_last_stage_ms(self._h, ms, C.c_uint32(9)), "last_stage_ms")
        return dict(zip(self.STAGES, (float(v) for v in ms)))

    def launch_count(self) -> int:
        return int(lib().lidar_b200_launch_count(self._h))

    def graph_launch_count(self) -> int:
        """single-frame batch_run calls replayed as one CUDA-graph launch"""
        return int(lib().lidar_b200_graph_launch_count(self._h))

    def last_replay_stats(self, capacity: int = 65536) -> np.ndarray:
        """(jobs, 8) uint32: frame, members, kilo-cycles, rounds, direct rounds, entries taken, seeds, candidates."""
        out = np.zeros((capacity, 8), np.uint32)
        n = C.c_uint32(0)
        self._check(lib().lidar_b200_last_replay_stats(self._h, _ptr(out, C.c_uint32), C.c_uint32(capacity), C.byref(n)),
                    "last_replay_stats")
        return out[: min(n.value, capacity)]

    def region_begin(self):
        """CUDA event on the context's stream: start of a device-timed region (see region_end_ms)."""
        self._check(lib().lidar_b200_region_begin(self._h), "region_begin")

    def region_end_ms(self) -> float:
        """Device milliseconds since region_begin() once everything enqueued in between has finished."""
        ms = C.c_float(0)
        self._check(lib().lidar_b200_region_end_ms(self._h, C.byref(ms)), "region_end_ms")
        return float(ms.value)

    def last_run_ms(self) -> float:
        ms = C.c_float(0)
        self._check(lib().lidar_b200_last_run_ms(self._h, C.byref(ms)), "last_run_ms")
        return float(ms.value)


def equal_chunks(n_frames: int, chunk_frames: int):
    """[(first, end), ...]: the fewest chunks of at most `chunk_frames` frames, sizes differing by at most one."""
    n_chunks = max(1, -(-n_frames // max(1, chunk_frames)))
    size, extra = divmod(n_frames, n_chunks)
    chunks, a = [], 0
    for k in range(n_chunks):
        b = a + size + (1 if k < extra else 0)
        if b > a:
            chunks.append((a, b))
        a = b
    return chunks


class FramePipeline:
    """Throughput path: a job of independent frames is cut into chunks that rotate through `depth`
    contexts (lidar_b200_pipe_*), so uploads, kernels and downloads of neighbouring chunks overlap.
    Results are written into one page-locked arena owned by the pipeline and returned as views: they
    stay valid until the next process() call."""

    def __init__(self, device: int = 0, depth: int = 3, chunk_frames: int = 22, max_points_per_chunk: int = 0,
                 gpus_sharing_host: int = 1):
        """gpus_sharing_host: how many GPUs of this host run a pipeline at the same time (the local world size); it
        selects the result fetch mode (lidar_b200_pipe_set_host_sharing)."""
        self._h = C.c_void_p()
        self.depth, self.chunk_frames = depth, chunk_frames
        rc = lib().lidar_b200_pipe_create(C.c_int(device), C.c_uint32(depth),
                                          C.c_uint32(max_points_per_chunk or 130_000 * chunk_frames),
                                          C.c_uint32(chunk_frames), C.byref(self._h))
        if rc != 0:
            raise LidarB200Error(f"lidar_b200_pipe_create failed (status {rc}): CUDA device {device} unavailable; "
                                 "there is no CPU fallback")
        self._check(lib().lidar_b200_pipe_set_host_sharing(self._h, C.c_uint32(max(1, gpus_sharing_host))), "pipe_set_host_sharing")
        self._arenas = {}
        self.h2d_bytes = self.d2h_bytes = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().lidar_b200_pipe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = lib().lidar_b200_pipe_last_error(self._h)
            raise LidarB200Error(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")

    def seg_configure(self, cfg: SegmentationConfiguration):
        self._check(lib().lidar_b200_pipe_seg_configure(self._h, C.byref(cfg)), "pipe_seg_configure")

    def clu_configure(self, cfg: ClusteringConfiguration):
        self._check(lib().lidar_b200_pipe_clu_configure(self._h, C.byref(cfg)), "pipe_clu_configure")

    def launch_count(self) -> int:
        return int(lib().lidar_b200_pipe_launch_count(self._h))

    def fetch_mode(self) -> int:
        """0 = slot-size copies by the copy engines, 4 = exact sizes written by a kernel (lidar_b200_pipe_fetch_mode)."""
        return int(lib().lidar_b200_pipe_fetch_mode(self._h))

    def submit(self, frames, want_ground_idx: bool = True, arena: int = 0):
        """Enqueues a job (all its chunks) without waiting; returns a handle for results(). Jobs submitted
        back to back keep the copy engines and the SMs busy across job boundaries; a job's results are
        complete after drain() (or once `depth` later chunks have been submitted). Use a different
        `arena` index for consecutive jobs whose results must stay readable meanwhile."""
        frames = [_as_points(f) for f in frames]
        nf = len(frames)
        strides = {f.shape[1] * 4 for f in frames} or {16}
        if len(strides) != 1:
            raise ValueError("all frames of a job must share one point stride")
        stride = strides.pop()
        counts = np.array([f.shape[0] for f in frames], np.uint32)
        padded = (counts.astype(np.int64) + 31) & ~31
        # chunks of equal size, at most chunk_frames frames each (a 154-frame job with chunk_frames = 51 is cut into
        # 39 + 39 + 38 + 38, not 51 + 51 + 51 + 1: a short last chunk leaves the pipeline's contexts idle)
        chunks = equal_chunks(nf, self.chunk_frames)
        total = int(padded.sum())
        ar = self._arenas.get(arena)
        if ar is None or ar[0].size < max(total, 1) or ar[4].shape[1] < max(nf, 1):
            ar = (pinned_empty(max(total, 1), np.uint32), pinned_empty(max(total, 1), np.uint32),
                  pinned_empty(max(total, 1), np.uint32), pinned_empty(max(total, 1), np.int32),
                  pinned_empty((4, max(nf, 1)), np.uint32))
            self._arenas[arena] = ar
        seg, gidx, oidx, clab, meta = ar
        u32, i32 = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)

        def at(a, pos, t):
            return C.cast(C.c_void_p(a.ctypes.data + 4 * pos), t)

        base = 0
        bases = []
        for a, b in chunks:
            n = b - a
            ptrs = (C.c_void_p * n)(*[f.ctypes.data for f in frames[a:b]])
            mrow = lambda r: C.cast(C.c_void_p(meta.ctypes.data + 4 * (r * meta.shape[1] + a)), u32)  # noqa: E731
            self._check(lib().lidar_b200_pipe_submit(
                self._h, C.c_uint32(n), ptrs, _ptr(counts[a:b], C.c_uint32), C.c_uint32(stride), mrow(0),
                at(seg, base, u32), at(gidx, base, u32) if want_ground_idx else None, mrow(1), at(oidx, base, u32),
                mrow(2), at(clab, base, i32), mrow(3)), "pipe_submit")
            bases.append(base)
            base += int(padded[a:b].sum())
        self.h2d_bytes = int(counts.astype(np.int64).sum()) * 16 + 16 * nf
        # default fetch mode: the whole slot range of the four result arrays; with LIDAR_B200_FETCH_MODE >= 2 the library
        # copies exact sizes (4 B per point, 4 B per ground point, 8 B per obstacle point) and results() corrects this
        self.d2h_bytes = (4 if want_ground_idx else 3) * total * 4 + 12 * nf + 4 * len(chunks)
        return dict(frames=frames, counts=counts, chunks=chunks, bases=bases, arena=arena, want_ground_idx=want_ground_idx)

    def drain(self):
        self._check(lib().lidar_b200_pipe_drain(self._h), "pipe_drain")

    def results(self, job):
        """Per-frame result views of a drained job (valid until its arena is reused)."""
        seg, gidx, oidx, clab, meta = self._arenas[job["arena"]]
        counts, want = job["counts"], job["want_ground_idx"]
        out = []
        for (a, b), cb in zip(job["chunks"], job["bases"]):
            for f in range(a, b):
                o, n = cb + int(meta[0, f]), int(counts[f])
                ng, no = int(meta[1, f]), int(meta[2, f])
                out.append(dict(seg_labels=seg[o:o + n], ground_idx=gidx[o:o + ng] if want else None,
                                obstacle_idx=oidx[o:o + no], cluster_labels=clab[o:o + no], n_clusters=int(meta[3, f])))
        nf = len(out)
        if self.fetch_mode() >= 2:  # exact-size results (see api.cu)
            self.d2h_bytes = (4 * int(counts.astype(np.int64).sum())
                              + (4 * int(meta[1, :nf].astype(np.int64).sum()) if want else 0)
                              + 8 * int(meta[2, :nf].astype(np.int64).sum()) + 12 * nf + 4 * len(job["chunks"]))
        return out

    def process(self, frames, want_ground_idx: bool = True):
        job = self.submit(frames, want_ground_idx)
        self.drain()
        return self.results(job)


class Segmenter:
    """Mirror of lidar_processing::Segmenter (reference src/segmentation.hpp:58-70)."""

    def __init__(self, device: int = 0, context: Context | None = None):
        self._ctx = context or Context(device)
        self._cfg = SegmentationConfiguration()

    def update_configuration(self, configuration: SegmentationConfiguration):
        self._ctx.seg_configure(configuration)
        self._cfg = configuration

    def reserve_memory(self, number_of_points: int = 200_000):
        self._ctx.reserve(number_of_points)

    def segment(self, cloud_in, labels: np.ndarray | None = None):
        """Returns (labels, ground_cloud, obstacle_cloud). `labels` (uint32, len >= N) is updated in
        place when given: only classified points are written, like the reference's resize()."""
        pts = _as_points(cloud_in)
        labels, gi, oi = self._ctx.segment(pts, labels)
        return labels, pts[gi], pts[oi]

    def segment_indices(self, cloud_in, labels: np.ndarray | None = None):
        return self._ctx.segment(_as_points(cloud_in), labels)


class Clusterer:
    """Mirror of lidar_processing::Clusterer (reference src/clustering.hpp:50-64)."""

    UNDEFINED = UNDEFINED
    INVALID = INVALID

    def __init__(self, device: int = 0, context: Context | None = None):
        self._ctx = context or Context(device)

    def update_configuration(self, configuration: ClusteringConfiguration):
        self._ctx.clu_configure(configuration)

    def reserve_memory(self, number_of_points: int = 200_000):
        self._ctx.reserve(number_of_points)

    def cluster(self, cloud_in) -> np.ndarray:
        return self._ctx.cluster(cloud_in)
