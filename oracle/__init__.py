"""TEST INFRASTRUCTURE ONLY — ctypes front end of the CPU oracle (oracle/oracle.cpp) and of the
unmodified reference Clusterer build (oracle/_ref/libref_cluster.so).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package; the
product path (lidar-processing_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = None
_REF = None

UNKNOWN, GROUND, OBSTACLE = 0, 1, 2
UNDEFINED = np.iinfo(np.int32).min
INVALID = -1


class SegCfg(C.Structure):
    """reference: src/segmentation.hpp:48-56"""

    _fields_ = [
        ("sensor_height_m", C.c_float),
        ("orthogonal_distance_threshold", C.c_float),
        ("initial_seed_threshold", C.c_float),
        ("number_of_iterations", C.c_uint32),
        ("number_of_planar_partitions", C.c_uint32),
        ("number_of_lower_point_representatives", C.c_uint32),
    ]


class CluCfg(C.Structure):
    """reference: src/clustering.hpp:42-48"""

    _fields_ = [
        ("distance_squared", C.c_float),
        ("cluster_quality", C.c_float),
        ("min_cluster_size", C.c_uint32),
        ("max_cluster_size", C.c_uint32),
    ]


def build(force: bool = False) -> None:
    """Compile liboracle.so, and oracle/_ref when /root/reference is present (else keep prebuilt)."""
    lib = HERE / "liboracle.so"
    lib_deps = [HERE / "oracle.cpp", HERE / "hull_oracle.cpp", HERE / "oracle.h", HERE / "Makefile"]
    ref = HERE / "_ref" / "libref_node.so"  # the last target of `make ref`
    ref_deps = lib_deps + [HERE / "ref_wrap.cpp", HERE / "ref_hull_wrap.cpp", HERE / "ref_seg_wrap.cpp",
                           HERE / "eigen_shim" / "Eigen" / "Dense", HERE / "ref_node_wrap.cpp",
                           HERE / "ros_shim" / "rclcpp" / "rclcpp.hpp"]
    need = force or not lib.exists() or lib.stat().st_mtime < max(s.stat().st_mtime for s in lib_deps)
    have_reference = Path("/root/reference/src/clustering.cpp").exists()
    if have_reference and (force or not ref.exists() or ref.stat().st_mtime < max(s.stat().st_mtime for s in ref_deps)):
        need = True
    if need:
        subprocess.run(["make", "-s", "-C", str(HERE), "all"], check=True)


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(str(HERE / "liboracle.so"))
        _LIB.oracle_fnv1a64.restype = C.c_uint64
        _LIB.oracle_canonicalise.restype = C.c_uint32
    return _LIB


def ref_available() -> bool:
    return (HERE / "_ref" / "libref_cluster.so").exists()


def ref() -> C.CDLL:
    global _REF
    if _REF is None:
        build()
        _REF = C.CDLL(str(HERE / "_ref" / "libref_cluster.so"))
    return _REF


def _f32(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] >= 3
    return a


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def default_seg_cfg(**kw) -> SegCfg:
    cfg = SegCfg()
    lib().oracle_seg_cfg_default(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def default_clu_cfg(**kw) -> CluCfg:
    cfg = CluCfg()
    lib().oracle_clu_cfg_default(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def segment(points, cfg: SegCfg | None = None, tie_mode: int = 1, labels_in=None):
    """Segmenter::segment. Returns dict(labels, ground_idx, obstacle_idx, planes, status)."""
    pts = _f32(points)
    n = pts.shape[0]
    cfg = cfg or default_seg_cfg()
    labels = np.zeros(n, np.uint32) if labels_in is None else np.ascontiguousarray(labels_in, np.uint32).copy()
    g = np.zeros(max(n, 1), np.uint32)
    o = np.zeros(max(n, 1), np.uint32)
    ng, no = C.c_uint32(0), C.c_uint32(0)
    P, it = cfg.number_of_planar_partitions, cfg.number_of_iterations
    planes = np.zeros((max(P, 1), max(it, 1), 4), np.float32)
    status = np.zeros(max(P, 1), np.int32)
    rc = lib().oracle_segment(_p(pts, C.c_float), C.c_uint32(n), C.c_uint32(pts.shape[1]), C.byref(cfg),
                              C.c_int(tie_mode), _p(labels, C.c_uint32), _p(g, C.c_uint32), C.byref(ng),
                              _p(o, C.c_uint32), C.byref(no), _p(planes, C.c_float), _p(status, C.c_int32))
    if rc != 0:
        raise ValueError("oracle_segment: bad configuration")
    return dict(labels=labels, ground_idx=g[: ng.value].copy(), obstacle_idx=o[: no.value].copy(),
                planes=planes[:P, :it], status=status[:P])


_REF_SEG = None


def ref_segment_available() -> bool:
    return (HERE / "_ref" / "libref_segment.so").exists()


def ref_seg() -> C.CDLL:
    """The UNMODIFIED reference Segmenter (src/segmentation.cpp) compiled against the PCL and Eigen stand-ins."""
    global _REF_SEG
    if _REF_SEG is None:
        build()
        _REF_SEG = C.CDLL(str(HERE / "_ref" / "libref_segment.so"))
    return _REF_SEG


def ref_segment(points, cfg: SegCfg | None = None, labels_in=None):
    """lidar_processing::Segmenter::segment of the reference itself. Returns dict(labels, ground_idx, obstacle_idx)."""
    pts = _f32(points)
    n = pts.shape[0]
    cfg = cfg or default_seg_cfg()
    labels = np.zeros(n, np.uint32) if labels_in is None else np.ascontiguousarray(labels_in, np.uint32).copy()
    g = np.zeros(max(n, 1), np.uint32)
    o = np.zeros(max(n, 1), np.uint32)
    ng, no = C.c_uint32(0), C.c_uint32(0)
    rc = ref_seg().ref_segment(_p(pts, C.c_float), C.c_uint32(n), C.c_uint32(pts.shape[1]), C.byref(cfg),
                               _p(labels, C.c_uint32), _p(g, C.c_uint32), C.byref(ng), _p(o, C.c_uint32), C.byref(no))
    if rc != 0:
        raise ValueError(f"ref_segment: {rc}")
    return dict(labels=labels, ground_idx=g[: ng.value].copy(), obstacle_idx=o[: no.value].copy())


def ref_segment_pair(points_a, points_b, cfg: SegCfg | None = None) -> np.ndarray:
    """Two frames through one long-lived reference Segmenter and one labels vector (processor.cpp:129-131,150)."""
    a, b = _f32(points_a), _f32(points_b)
    assert a.shape[1] == b.shape[1]
    cfg = cfg or default_seg_cfg()
    out = np.zeros(max(a.shape[0], b.shape[0], 1), np.uint32)
    n = ref_seg().ref_segment_pair(_p(a, C.c_float), C.c_uint32(a.shape[0]), _p(b, C.c_float), C.c_uint32(b.shape[0]),
                                   C.c_uint32(a.shape[1]), C.byref(cfg), _p(out, C.c_uint32))
    if n < 0:
        raise ValueError(f"ref_segment_pair: {n}")
    return out[:n].copy()


_REF_NODE = None


def ref_node_available() -> bool:
    return (HERE / "_ref" / "libref_node.so").exists()


def ref_node_run(points, rand_seed: int = 1):
    """One frame through the UNMODIFIED reference node (src/processor.cpp: Processor::process) built against the ROS 2 /
    PCL / Eigen stand-ins. Returns what it publishes: dict(ground, obstacle (n, 32) uint8 pcl::PointXYZRGBL records or
    None, clustered (n, 32) uint8 pcl::PointXYZRGB records or None, marker_ids, markers = list of (k, 3) float64)."""
    global _REF_NODE
    if _REF_NODE is None:
        build()
        _REF_NODE = C.CDLL(str(HERE / "_ref" / "libref_node.so"))
    pts = _f32(points)
    rc = _REF_NODE.ref_node_run(_p(pts, C.c_float), C.c_uint32(pts.shape[0]), C.c_uint32(pts.shape[1]), C.c_uint(rand_seed))
    if rc != 0:
        raise RuntimeError("the reference node reported an exception")
    sizes = np.zeros(9, np.uint64)
    _REF_NODE.ref_node_sizes(_p(sizes, C.c_uint64))
    ground = np.zeros(int(sizes[0]), np.uint8)
    obstacle = np.zeros(int(sizes[1]), np.uint8)
    clustered = np.zeros(int(sizes[2]), np.uint8)
    msz = np.zeros(max(int(sizes[3]), 1), np.uint32)
    mid = np.zeros(max(int(sizes[3]), 1), np.int32)
    mxyz = np.zeros((max(int(sizes[4]), 1), 3), np.float64)
    _REF_NODE.ref_node_fetch(_p(ground, C.c_uint8), _p(obstacle, C.c_uint8), _p(clustered, C.c_uint8), _p(msz, C.c_uint32),
                             _p(mid, C.c_int32), _p(mxyz, C.c_double))
    topics = int(sizes[8])
    assert all(int(sizes[k]) in (0, 32) for k in (5, 6, 7))  # point_step of every published cloud
    markers, at = [], 0
    for k in range(int(sizes[3])):
        markers.append(mxyz[at:at + int(msz[k])].copy())
        at += int(msz[k])
    return dict(ground=ground.reshape(-1, 32) if topics & 1 else None, obstacle=obstacle.reshape(-1, 32) if topics & 2 else None,
                clustered=clustered.reshape(-1, 32) if topics & 4 else None, marker_ids=mid[: int(sizes[3])].copy(),
                markers=markers if topics & 8 else None)


def libc_rand_colors(n_clusters: int, rand_seed: int = 1) -> np.ndarray:
    """The colour words r << 16 | g << 8 | b the node draws with std::rand() % 256 after std::srand(rand_seed)
    (reference src/conversions.cpp:48-50), from the same C library."""
    libc = C.CDLL(None)
    libc.srand(C.c_uint(rand_seed))
    out = np.zeros(n_clusters, np.uint32)
    for k in range(n_clusters):
        r, g, b = libc.rand() % 256, libc.rand() % 256, libc.rand() % 256
        out[k] = (r << 16) | (g << 8) | b
    return out


def jacobi_svd3(a):
    a = np.ascontiguousarray(a, np.float32).reshape(9)
    v = np.zeros(9, np.float32)
    sv = np.zeros(3, np.float32)
    sweeps = lib().oracle_jacobi_svd3(_p(a, C.c_float), _p(v, C.c_float), _p(sv, C.c_float))
    return v.reshape(3, 3), sv, sweeps


def cluster(points, cfg: CluCfg | None = None) -> np.ndarray:
    pts = _f32(points)
    m = pts.shape[0]
    cfg = cfg or default_clu_cfg()
    labels = np.zeros(max(m, 1), np.int32)
    lib().oracle_cluster(_p(pts, C.c_float), C.c_uint32(m), C.c_uint32(pts.shape[1]), C.byref(cfg),
                         _p(labels, C.c_int32))
    return labels[:m]


def kd_order(points, mode: int = 0) -> np.ndarray:
    pts = _f32(points)
    m = pts.shape[0]
    order = np.zeros(max(m, 1), np.uint32)
    lib().oracle_kd_order(_p(pts, C.c_float), C.c_uint32(m), C.c_uint32(pts.shape[1]), C.c_int(mode),
                          _p(order, C.c_uint32))
    return order[:m]


def kd_build(points, mode: int = 0) -> np.ndarray:
    pts = _f32(points)
    m = pts.shape[0]
    slots = np.zeros(max(m, 1), np.uint32)
    lib().oracle_kd_build(_p(pts, C.c_float), C.c_uint32(m), C.c_uint32(pts.shape[1]), C.c_int(mode),
                          _p(slots, C.c_uint32))
    return slots[:m]


def kd_rank(points, mode: int = 0) -> np.ndarray:
    order = kd_order(points, mode)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size, dtype=np.uint32)
    return rank


def cc_roots(points, distance_squared: float = 0.18) -> np.ndarray:
    pts = _f32(points)
    m = pts.shape[0]
    root = np.zeros(max(m, 1), np.uint32)
    lib().oracle_cc(_p(pts, C.c_float), C.c_uint32(m), C.c_uint32(pts.shape[1]), C.c_float(distance_squared),
                    _p(root, C.c_uint32))
    return root[:m]


def cluster_model(points, rank=None, cfg: CluCfg | None = None):
    pts = _f32(points)
    m = pts.shape[0]
    cfg = cfg or default_clu_cfg()
    if rank is None:
        rank = kd_rank(pts, 0)
    rank = np.ascontiguousarray(rank, np.uint32)
    labels = np.zeros(max(m, 1), np.int32)
    stats = np.zeros(8, np.uint64)
    lib().oracle_cluster_model(_p(pts, C.c_float), C.c_uint32(m), C.c_uint32(pts.shape[1]), C.byref(cfg),
                               _p(rank, C.c_uint32), _p(labels, C.c_int32), _p(stats, C.c_uint64))
    names = ["seeds", "expansions", "pushes", "max_queue", "components", "max_component", "max_comp_expansions",
             "hits"]
    return labels[:m], dict(zip(names, (int(x) for x in stats)))


def canonicalise(labels) -> np.ndarray:
    lab = np.ascontiguousarray(labels, np.int32)
    out = np.zeros(max(lab.size, 1), np.int32)
    lib().oracle_canonicalise(_p(lab, C.c_int32), C.c_uint32(lab.size), _p(out, C.c_int32))
    return out[: lab.size]


def fnv1a64(arr) -> int:
    a = np.ascontiguousarray(arr)
    return int(lib().oracle_fnv1a64(a.ctypes.data_as(C.c_void_p), C.c_uint64(a.nbytes)))


# ---- unmodified reference build (oracle/_ref) ------------------------------------------------

def ref_cluster(points, cfg: CluCfg | None = None) -> np.ndarray:
    pts = _f32(points)
    m = pts.shape[0]
    cfg = cfg or default_clu_cfg()
    labels = np.zeros(max(m, 1), np.int32)
    ref().ref_cluster(_p(pts, C.c_float), C.c_uint32(m), C.c_uint32(pts.shape[1]), C.c_float(cfg.distance_squared),
                      C.c_float(cfg.cluster_quality), C.c_uint32(cfg.min_cluster_size),
                      C.c_uint32(cfg.max_cluster_size), _p(labels, C.c_int32))
    return labels[:m]


def ref_cluster_timed(points, repeats: int = 3):
    pts = _f32(points)
    m = pts.shape[0]
    labels = np.zeros(max(m, 1), np.int32)
    best, mean = C.c_double(0), C.c_double(0)
    ref().ref_cluster_timed(_p(pts, C.c_float), C.c_uint32(m), C.c_uint32(pts.shape[1]), C.c_uint32(repeats),
                            _p(labels, C.c_int32), C.byref(best), C.byref(mean))
    return labels[:m], best.value, mean.value


def ref_kd_order(points) -> np.ndarray:
    pts = _f32(points)
    m = pts.shape[0]
    order = np.zeros(max(m, 1), np.uint32)
    rc = ref().ref_kd_preorder(_p(pts, C.c_float), C.c_uint32(m), C.c_uint32(pts.shape[1]), _p(order, C.c_uint32))
    assert rc == 0
    return order[:m]


def ref_radius_search(points, queries, radius_sqr: float = 0.18, capacity: int | None = None):
    pts = _f32(points)
    m = pts.shape[0]
    q = np.ascontiguousarray(queries, np.uint32)
    capacity = capacity or max(1, 64 * 1024 * 1024 // 8)
    offs = np.zeros(q.size + 1, np.uint32)
    idx = np.zeros(capacity, np.uint32)
    d2 = np.zeros(capacity, np.float32)
    rc = ref().ref_radius_search(_p(pts, C.c_float), C.c_uint32(m), C.c_uint32(pts.shape[1]), _p(q, C.c_uint32),
                                 C.c_uint32(q.size), C.c_float(radius_sqr), _p(offs, C.c_uint32),
                                 _p(idx, C.c_uint32), _p(d2, C.c_float), C.c_uint32(capacity))
    assert rc == 0, rc
    return offs, idx[: offs[-1]].copy(), d2[: offs[-1]].copy()


def ref_pipeline_run(frames, nthreads: int = 1):
    """CPU baseline: restated Segmenter + unmodified reference Clusterer on `nthreads` threads.
    Returns dict(per_frame_ms, wall_s, n_obstacle, n_clusters)."""
    frames = [_f32(f) for f in frames]
    nf = len(frames)
    strides = {f.shape[1] for f in frames}
    assert len(strides) == 1
    ptrs = (C.POINTER(C.c_float) * nf)(*[_p(f, C.c_float) for f in frames])
    counts = np.array([f.shape[0] for f in frames], np.uint32)
    ms = np.zeros(nf, np.float64)
    wall = C.c_double(0)
    nobs = np.zeros(nf, np.uint32)
    ncl = np.zeros(nf, np.uint32)
    ref().ref_pipeline_run(ptrs, _p(counts, C.c_uint32), C.c_uint32(nf), C.c_uint32(strides.pop()), C.c_uint32(nthreads),
                           _p(ms, C.c_double), C.byref(wall), _p(nobs, C.c_uint32), _p(ncl, C.c_uint32))
    return dict(per_frame_ms=ms, wall_s=wall.value, n_obstacle=nobs, n_clusters=ncl)


def ref_pipeline_passes(frames, nthreads: int = 1, n_passes: int = 1):
    """Like ref_pipeline_run over `n_passes` passes of the frame list with long-lived worker threads and Clusterers
    (nothing is constructed inside a timed pass). Returns dict(pass_wall_s[n_passes], per_frame_ms (last pass), ...)."""
    frames = [_f32(f) for f in frames]
    nf = len(frames)
    strides = {f.shape[1] for f in frames}
    assert len(strides) == 1
    ptrs = (C.POINTER(C.c_float) * nf)(*[_p(f, C.c_float) for f in frames])
    counts = np.array([f.shape[0] for f in frames], np.uint32)
    ms = np.zeros(nf, np.float64)
    walls = np.zeros(max(n_passes, 1), np.float64)
    nobs = np.zeros(nf, np.uint32)
    ncl = np.zeros(nf, np.uint32)
    ref().ref_pipeline_passes(ptrs, _p(counts, C.c_uint32), C.c_uint32(nf), C.c_uint32(strides.pop()), C.c_uint32(nthreads),
                              C.c_uint32(n_passes), _p(walls, C.c_double), _p(ms, C.c_double), _p(nobs, C.c_uint32),
                              _p(ncl, C.c_uint32))
    return dict(pass_wall_s=walls[:n_passes], per_frame_ms=ms, n_obstacle=nobs, n_clusters=ncl)


# ---- PCD v0.7 "DATA binary" reader (dataloader.cpp:139 uses pcl::io::loadPCDFile) ------------

def read_pcd(path) -> np.ndarray:
    """Returns an (N, 4) float32 array x, y, z, intensity."""
    with open(path, "rb") as f:
        raw = f.read()
    pos = 0
    npts = None
    fields = None
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if line.startswith("FIELDS"):
            fields = line.split()[1:]
        elif line.startswith("POINTS"):
            npts = int(line.split()[1])
        elif line.startswith("DATA"):
            assert line.split()[1] == "binary", "only DATA binary is supported"
            break
    assert fields == ["x", "y", "z", "intensity"] and npts is not None
    return np.frombuffer(raw, dtype=np.float32, count=npts * 4, offset=pos).reshape(npts, 4).copy()


REFERENCE_DATA = Path(os.environ.get("LIDAR_B200_REFERENCE_DATA", "/root/reference/data"))


def reference_frame_paths():
    if not REFERENCE_DATA.is_dir():
        return []
    return sorted(REFERENCE_DATA.glob("*.pcd"))


def split_clusters(obstacle_points, labels):
    """CPU restatement of the per-cluster split in Processor::process (reference src/processor.cpp:180-200):
    clustered_obstacle_cloud.resize(max_label + 1); for i ascending: UNDEFINED -> runtime_error, INVALID
    skipped, else clustered_obstacle_cloud[label].emplace_back(x, y, z); empty clouds erased.
    Returns a list of (points[n_k,3] float32, source indices) per remaining cluster, in label order."""
    labels = np.asarray(labels)
    pts = np.asarray(obstacle_points, np.float32)[:, :3]
    if labels.size == 0:
        return []  # (the reference's max_element on an empty vector is UB in the caller)
    if np.any(labels == UNDEFINED):
        raise RuntimeError("Undefined label found (clustering)")
    clouds = [[] for _ in range(int(labels.max()) + 1)]
    for i, l in enumerate(labels.tolist()):
        if l != INVALID:
            clouds[l].append(i)
    out = []
    for idx in clouds:
        if idx:  # std::remove_if(cloud.empty())
            ii = np.asarray(idx, np.int64)
            out.append((pts[ii], ii))
    return out


_REF_HULL = None


def ref_hull_available() -> bool:
    return (HERE / "_ref" / "libref_hull.so").exists()


def _clusters_csr(clusters):
    """list of (n_k, >=2) arrays -> (points[n,4] float32, offsets[K+1] uint32)"""
    offsets = np.zeros(len(clusters) + 1, np.uint32)
    if clusters:
        offsets[1:] = np.cumsum([len(c) for c in clusters])
    pts = np.zeros((max(int(offsets[-1]), 1), 4), np.float32)
    at = 0
    for c in clusters:
        c = np.asarray(c, np.float32)
        pts[at:at + len(c), :min(c.shape[1], 3)] = c[:, :3] if c.size else 0
        at += len(c)
    return pts, offsets


def convex_outlines(clusters, mode: int = 0):
    """Restated findOrderedConvexOutlines (mode 0) / convex branch of findOrderedConcaveOutlines (mode 1),
    oracle/hull_oracle.cpp. Returns one (xy[h,2] float32, local_idx[h]) per cluster (h = 0: dropped / host; None: the
    reference's Jarvis march would never close on this cluster, i.e. the reference hangs)."""
    pts, offsets = _clusters_csr(clusters)
    k = len(clusters)
    sizes = np.zeros(max(k, 1), np.uint32)
    idx = np.zeros(max(int(offsets[-1]) * 2, 1), np.uint32)
    f = lib().oracle_convex_outlines
    f.restype = C.c_longlong
    n = f(_p(pts, C.c_float), _p(offsets, C.c_uint32), C.c_uint32(k), C.c_uint32(4), C.c_int(mode),
          _p(sizes, C.c_uint32), _p(idx, C.c_uint32), C.c_longlong(idx.size))
    assert n >= 0
    out, at = [], 0
    for c in range(k):
        h = int(sizes[c])
        if h == 0xFFFFFFFF:  # the reference's Jarvis march does not terminate on this cluster
            out.append(None)
            continue
        li = idx[at:at + h].astype(np.int64)
        out.append((pts[int(offsets[c]) + li][:, :2].copy(), li))
        at += h
    return out


def ref_outlines(clusters, mode: int = 0):
    """UNMODIFIED reference findOrderedConvexOutlines (mode 0) / findOrderedConcaveOutlines (mode 1).
    Returns one xy[h,2] float32 array per cluster (h = 0: the reference dropped the outline; None: it threw)."""
    global _REF_HULL
    if _REF_HULL is None:
        build()
        _REF_HULL = C.CDLL(str(HERE / "_ref" / "libref_hull.so"))
        _REF_HULL.ref_outlines.restype = C.c_longlong
    pts, offsets = _clusters_csr(clusters)
    k = len(clusters)
    sizes = np.zeros(max(k, 1), np.uint32)
    xy = np.zeros((max(int(offsets[-1]) * 2, 1), 2), np.float32)
    n = _REF_HULL.ref_outlines(_p(pts, C.c_float), _p(offsets, C.c_uint32), C.c_uint32(k), C.c_uint32(4), C.c_int(mode),
                               _p(sizes, C.c_uint32), _p(xy, C.c_float), C.c_longlong(xy.shape[0]))
    assert n >= 0
    out, at = [], 0
    for c in range(k):
        h = int(sizes[c])
        if h == 0xFFFFFFFF:  # the reference threw (delaunator on degenerate input)
            out.append(None)
            continue
        out.append(xy[at:at + h].copy())
        at += h
    return out


def colorize(clusters, cluster_rgb) -> np.ndarray:
    """Restated convertClusteredCloudToColorizedCloud (reference src/conversions.cpp:32-60) + the byte layout
    convertPCLToPointCloud2 copies out (conversions.cpp:139-162): (n, 32) uint8 pcl::PointXYZRGB records, cluster
    after cluster. cluster_rgb[k] = r << 16 | g << 8 | b (the reference draws r, g, b with std::rand() % 256).
    Parity unpinned at the PCL boundary: PCL is absent here; layout per pcl/impl/point_types.hpp (PCL_ADD_POINT4D,
    PCL_ADD_RGB: b, g, r, a bytes, a = 255, 16-byte aligned, 32 bytes)."""
    n = sum(len(c) for c in clusters)
    out = np.zeros((n, 32), np.uint8)
    at = 0
    for c, word in zip(clusters, cluster_rgb):
        c = np.asarray(c, np.float32)
        m = len(c)
        rec = out[at:at + m]
        xyz1 = np.ones((m, 4), np.float32)
        xyz1[:, :3] = c[:, :3]
        rec[:, :16] = xyz1.view(np.uint8).reshape(m, 16)
        w = int(word)
        rec[:, 16] = w & 0xFF          # b
        rec[:, 17] = (w >> 8) & 0xFF   # g
        rec[:, 18] = (w >> 16) & 0xFF  # r
        rec[:, 19] = 255               # a
        at += m
    return out


def marker_points(outlines):
    """Restated point list of convertPointXYZTypeToMarkerArray (reference src/conversions.hpp:72-120): per non-empty
    outline (h, 2) -> (h + 1, 3) float64 {x, y, 0} with the first vertex appended (loop closure); empty -> None."""
    res = []
    for o in outlines:
        o = np.asarray(o, np.float32).reshape(-1, 2)
        if o.shape[0] == 0:
            res.append(None)
            continue
        p = np.zeros((o.shape[0] + 1, 3), np.float64)
        p[:-1, :2] = o.astype(np.float64)
        p[-1] = p[0]
        res.append(p)
    return res
