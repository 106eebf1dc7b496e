"""CPU tests of the product's host-compilable pieces and of the C-ABI library surface."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import oracle as O

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def hc(pkg):
    import __graft_entry__ as ge

    return C.CDLL(str(ge.build_host_checks()))


def _nodes(pts):
    a = np.zeros((len(pts), 4), np.float32)
    a[:, :3] = pts[:, :3]
    a[:, 3] = np.arange(len(pts), dtype=np.uint32).view(np.float32)
    return a


def _run(hc, a, first, nth, last, axis, mode, nw=8, cut=48):
    b = a.copy()
    hc.hc_nth_element(b.ctypes.data_as(C.c_void_p), first, nth, last, axis, mode, nw, cut)
    return b.view(np.uint32)


def test_nth_element_emulation_matches_libstdcxx(hc):
    """Product kd_select.h (sequential) and the lane-level model of the cooperative partition must
    leave exactly the permutation std::nth_element leaves, ties included."""
    rng = np.random.default_rng(1)
    for trial in range(1500):
        n = int(rng.integers(1, 300)) if trial % 3 else int(rng.integers(300, 5000))
        nvals = int(rng.choice([1, 2, 3, 5, 20, 200, 100000]))
        a = _nodes(rng.integers(0, nvals, size=(n, 3)).astype(np.float32) / 8)
        nth = n // 2 if trial % 2 else int(rng.integers(0, n))
        axis = int(rng.integers(0, 3))
        lo = int(rng.integers(0, max(1, n // 4))) if trial % 5 == 0 else 0
        nth = max(nth, lo)
        want = _run(hc, a, lo, nth, n, axis, 0)
        assert np.array_equal(_run(hc, a, lo, nth, n, axis, 1), want)
        assert np.array_equal(_run(hc, a, lo, nth, n, axis, 2, int(rng.choice([1, 8, 32])), int(rng.choice([3, 16, 48]))), want)


def test_kd_preorder_rank_of_product_headers(hc, golden_frames):
    pts = golden_frames[0]
    obs = pts[O.segment(pts, tie_mode=1)["obstacle_idx"]]
    want = O.kd_rank(obs, 0)
    for mode, nw, cut in ((1, 1, 3), (2, 32, 48), (2, 8, 48)):
        a = _nodes(obs)
        rank = np.zeros(len(obs), np.uint32)
        hc.hc_kd_build(a.ctypes.data_as(C.c_void_p), len(obs), mode, nw, cut, rank.ctypes.data_as(C.c_void_p))
        assert np.array_equal(rank, want)


def test_kd_range_at_partitions_the_array(hc):
    for m in (1, 2, 3, 7, 64, 1000, 46851):
        for depth in (0, 1, 3, 6):
            covered = np.zeros(m, np.int32)
            # nodes at shallower depths + ranges at this depth tile [0, m)
            for d in range(depth + 1):
                for path in range(1 << d):
                    b, e = C.c_uint32(), C.c_uint32()
                    if hc.hc_range_at(m, d, path, C.byref(b), C.byref(e)):
                        if d == depth:
                            covered[b.value:e.value] += 1
                        else:
                            covered[b.value + (e.value - b.value) // 2] += 1
            assert np.all(covered == 1)


def test_product_jacobi_equals_oracle_restatement(hc):
    rng = np.random.default_rng(2)
    for _ in range(300):
        m = rng.normal(size=(40, 3)) * rng.uniform(0.01, 30, 3)
        a = np.cov(m.T).astype(np.float32).reshape(9).copy()
        v = np.zeros(9, np.float32)
        sv = np.zeros(3, np.float32)
        assert hc.hc_jacobi_svd3(a.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), sv.ctypes.data_as(C.c_void_p)) == 1
        ov, osv, _ = O.jacobi_svd3(a)
        assert np.array_equal(v.reshape(3, 3).view(np.uint32), ov.view(np.uint32))
        assert np.array_equal(sv.view(np.uint32), osv.view(np.uint32))
    bad = np.full(9, np.nan, np.float32)
    assert hc.hc_jacobi_svd3(bad.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), sv.ctypes.data_as(C.c_void_p)) == 0


def test_c_abi_library_loads_and_exports_every_declared_symbol(pkg):
    header = (ROOT / "include" / "lidar_b200.h").read_text()
    declared = set(re.findall(r"\b(lidar_b200_[a-z0-9_]+)\s*\(", header))
    declared -= {"lidar_b200_ctx", "lidar_b200_status"}
    assert declared == set(pkg.EXPORTED_SYMBOLS)
    lib = pkg.lib()
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert b"sm_100a" in lib.lidar_b200_version()
    cfg = pkg.SegmentationConfiguration()
    lib.lidar_b200_seg_cfg_default(C.byref(cfg))
    assert (cfg.number_of_iterations, cfg.number_of_planar_partitions, cfg.number_of_lower_point_representatives) == (3, 2, 5000)
    assert abs(cfg.sensor_height_m - 1.73) < 1e-6 and abs(cfg.initial_seed_threshold - 0.6) < 1e-6
    ccfg = pkg.ClusteringConfiguration()
    lib.lidar_b200_clu_cfg_default(C.byref(ccfg))
    assert abs(ccfg.distance_squared - 0.18) < 1e-7 and ccfg.min_cluster_size == 4 and ccfg.max_cluster_size == 0xFFFFFFFF


def test_no_cpu_fallback_without_a_gpu(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.LidarB200Error):
        pkg.Context(device=0)
    with pytest.raises(pkg.LidarB200Error):
        pkg.Segmenter()


def test_product_never_touches_the_oracle():
    for p in (ROOT / "lidar-processing_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"}:
            text = p.read_text()
            assert "oracle" not in text.replace("the oracle", "").lower() or p.name == "README.md", p


def test_pcd_reader_and_frame_cache_round_trip(tmp_path, golden_frames):
    from tools.pack_reference_frames import pack, unpack

    pts = golden_frames[0][:1000]
    path = tmp_path / "f.pcd"
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\n"
           "TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 1000\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 1000\nDATA binary\n")
    path.write_bytes(hdr.encode() + pts.tobytes())
    back = O.read_pcd(path)
    assert np.array_equal(back.view(np.uint32), pts.view(np.uint32))
    pack([path], tmp_path / "c.xz")
    assert np.array_equal(unpack(tmp_path / "c.xz")[0].view(np.uint32), pts.view(np.uint32))


def test_library_pcd_reader(pkg, tmp_path, golden_frames):
    """lidar_b200_pcd_read (host code of the product, reference src/dataloader.cpp:87-126,139) against the
    oracle's reader: binary and ascii, field order, extra fields, both output layouts, error reporting."""
    pts = golden_frames[0][:777]
    n = pts.shape[0]
    base = "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\n"
    tail = f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\n"
    # 1. the layout of data/*.pcd
    p1 = tmp_path / "a.pcd"
    p1.write_bytes((base + "FIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n" + tail + "DATA binary\n").encode()
                   + pts.tobytes())
    got = pkg.read_pcd(p1)
    assert np.array_equal(got.view(np.uint32), O.read_pcd(p1).view(np.uint32))
    wire = pkg.read_pcd(p1, stride_bytes=32)  # pcl::PointXYZI: x y z 1.0 | intensity 0 0 0
    assert wire.shape == (n, 8) and np.array_equal(wire[:, :3], pts[:, :3]) and np.all(wire[:, 3] == 1.0)
    assert np.array_equal(wire[:, 4], pts[:, 3]) and not wire[:, 5:].any()
    # 2. permuted fields with an extra 2-byte x 3 field in the middle, no intensity
    rec = np.zeros(n, dtype=[("z", "<f4"), ("ring", "<u2", 3), ("x", "<f4"), ("y", "<f4")])
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    p2 = tmp_path / "b.pcd"
    p2.write_bytes((base + "FIELDS z ring x y\nSIZE 4 2 4 4\nTYPE F U F F\nCOUNT 1 3 1 1\n" + tail + "DATA binary\n").encode()
                   + rec.tobytes())
    got = pkg.read_pcd(p2)
    assert np.array_equal(got[:, :3].view(np.uint32), pts[:, :3].view(np.uint32)) and not got[:, 3].any()
    # 3. ascii (9 significant digits round-trip float32 exactly)
    p3 = tmp_path / "c.pcd"
    rows = "\n".join(" ".join(f"{v:.9g}" for v in r) for r in pts.tolist())
    p3.write_text(base + "FIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n" + tail + "DATA ascii\n" + rows + "\n")
    assert np.array_equal(pkg.read_pcd(p3).view(np.uint32), pts.view(np.uint32))
    # 4. errors are reported, not swallowed
    p4 = tmp_path / "d.pcd"
    p4.write_bytes((base + "FIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n" + tail + "DATA binary\n").encode()
                   + pts.tobytes()[:-100])
    with pytest.raises(pkg.LidarB200Error, match="shorter"):
        pkg.read_pcd(p4)
    p5 = tmp_path / "e.pcd"
    p5.write_text(base + "FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\n" + tail + "DATA binary_compressed\n")
    with pytest.raises(pkg.LidarB200Error, match="unsupported DATA"):
        pkg.read_pcd(p5)
    with pytest.raises(pkg.LidarB200Error, match="cannot open"):
        pkg.read_pcd(tmp_path / "missing.pcd")


def test_c_abi_rejects_a_null_context_without_touching_a_device(pkg):
    """Argument validation of the entry points added for SURVEY 8f rows 1, 3 and 4 (no GPU needed: a NULL context is
    refused before any CUDA call)."""
    lib = pkg.lib()
    header = (ROOT / "include" / "lidar_b200.h").read_text()
    inv = int(re.search(r"LIDAR_B200_ERR_INVALID\s*=\s*(\d+)", header).group(1))
    assert re.search(r"LIDAR_B200_ERR_INPUT\s*=\s*%d\b" % pkg.ERR_INPUT, header)
    buf = (C.c_uint32 * 4)()
    fbuf = (C.c_float * 8)()
    dbuf = (C.c_double * 6)()
    assert lib.lidar_b200_batch_group_clusters(None) == inv
    assert lib.lidar_b200_batch_fetch_clusters(None, buf, buf, fbuf, buf) == inv
    assert lib.lidar_b200_batch_hull_outlines(None, C.c_uint32(pkg.HULL_CONVEX)) == inv
    assert lib.lidar_b200_batch_fetch_hulls(None, buf, buf, fbuf, buf) == inv
    assert lib.lidar_b200_batch_fetch_colorized(None, buf, C.c_uint64(1), fbuf) == inv
    assert lib.lidar_b200_batch_fetch_marker_points(None, buf, buf, dbuf) == inv
    assert lib.lidar_b200_pipe_set_host_sharing(None, C.c_uint32(8)) == inv
    assert lib.lidar_b200_pipe_fetch_mode(None) == -1


# ---- csrc/chi_shape.h: the libstdc++ re-enactments behind the concave outline ------------------------------------------
def test_introsort_reenactment_matches_std_sort(hc):
    """lb::chi_introsort_ids against std::sort(ids, by dist) of this container's libstdc++ on tie-heavy keys: with
    ties the permutation std::sort leaves behind is exactly what the reference's sweep order depends on
    (Concave-Hull/delaunator.cpp:339-341). Includes the depth-limit (heap sort) path."""
    rng = np.random.default_rng(3)
    cases = []
    for n in list(range(0, 40)) + [100, 257, 1000, 4096, 20000]:
        for distinct in (1, 2, 3, 7, max(1, n // 4), max(1, n)):
            cases.append(rng.integers(0, distinct, size=n).astype(np.float64))
    # median-of-three killers drive introsort into its heap-sort branch
    for n in (64, 1000, 5000):
        k = np.zeros(n)
        half = n // 2
        for i in range(half):
            k[i] = (i + 1) if i % 2 == 0 else (half + i + (1 if half % 2 == 0 else 0))
            k[half + i] = 2 * (i + 1)
        cases.append(k)
        cases.append(np.minimum(k, n // 3))
    cases.append(np.arange(3000, dtype=np.float64)[::-1].copy())
    cases.append(np.abs(np.arange(-1500, 1500, dtype=np.float64)))
    for dist in cases:
        n = dist.size
        a = np.arange(n, dtype=np.uint32)
        b = a.copy()
        hc.hc_sort_ids(a.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p), n, 0)
        hc.hc_sort_ids(b.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p), n, 1)
        assert np.array_equal(a, b), f"n = {n}, {np.unique(dist).size} distinct keys"


def test_heap_reenactment_matches_std_heap(hc):
    """lb::chi_heap_push / chi_heap_pop against std::push_heap / std::pop_heap on pairs compared by length only
    (Concave-Hull/concave_hull.hpp:91-94): with equal lengths the pop order depends on the sift paths."""
    rng = np.random.default_rng(4)
    hc.hc_heap_replay.restype = C.c_uint32
    for trial in range(60):
        n = int(rng.integers(1, 3000))
        ops = np.where(rng.random(n) < (0.35 if trial % 2 else 0.5), -1, 1).astype(np.int32)
        ops[: min(n, 20)] = 1
        lens = rng.integers(0, max(2, n // (1 + trial % 7)), size=n).astype(np.float64) * 0.05
        out = [np.zeros(n, np.uint32), np.zeros(n, np.uint32)]
        cnt = [hc.hc_heap_replay(ops.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), n, m,
                                 out[m].ctypes.data_as(C.c_void_p)) for m in (0, 1)]
        assert cnt[0] == cnt[1] and np.array_equal(out[0][:cnt[0]], out[1][:cnt[1]])


def test_concave_working_set_fits_its_slot(hc):
    """chi_shape.cuh places the working set of a cluster of n >= 20 points at 144 bytes per point of its CSR range."""
    hc.hc_chi_layout_bytes.restype = C.c_ulonglong
    for n in list(range(20, 3000)) + [4095, 4096, 4097, 65535, 65536, 1 << 20, (1 << 24) + 1, (1 << 31) - 1]:
        assert hc.hc_chi_layout_bytes(C.c_uint32(n)) <= 144 * n, n


def test_pipeline_chunks_are_equal_sized(pkg):
    """A job is cut into the fewest chunks of at most chunk_frames frames, sizes differing by at most one."""
    for nf in (0, 1, 7, 22, 64, 154, 256, 4096):
        for cf in (1, 11, 22, 51, 77, 154, 1000):
            ch = pkg.equal_chunks(nf, cf)
            assert sum(b - a for a, b in ch) == nf and all(0 < b - a <= cf for a, b in ch)
            assert [a for a, _ in ch] == [0] + [b for _, b in ch][:-1] if ch else nf == 0
            if ch:
                sizes = [b - a for a, b in ch]
                assert max(sizes) - min(sizes) <= 1 and len(ch) == -(-nf // cf)
    assert pkg.equal_chunks(154, 51) == [(0, 39), (39, 78), (78, 116), (116, 154)]
