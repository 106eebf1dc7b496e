"""Shared parity checks: the CUDA path (through the C ABI) against the CPU oracle."""
from __future__ import annotations

import numpy as np

import oracle as O

FLIP_FRACTION = 1e-3   # north-star tolerance: at most 0.1 % of points may flip ...
FLIP_BAND_M = 1e-4     # ... and only within 1e-4 m of the decision surface


def to_oracle_seg_cfg(cfg):
    return O.default_seg_cfg(**{n: getattr(cfg, n) for n, _ in cfg._fields_})


def to_oracle_clu_cfg(cfg):
    return O.default_clu_cfg(**{n: getattr(cfg, n) for n, _ in cfg._fields_})


def segment_ids(pts, partitions):
    """partition id of every point under the stable x order (-1 = in no partition)."""
    n = pts.shape[0]
    order = np.argsort(pts[:, 0], kind="stable")
    per = n // partitions
    seg = np.full(n, -1, np.int64)
    for s in range(partitions):
        seg[order[s * per:(s + 1) * per]] = s
    return seg


def _margins(pts, planes_last, seg, thr):
    """signed distance of every point to the decision surface of its partition's last plane (float64)"""
    out = np.full(pts.shape[0], np.nan)
    X = pts[:, :3].astype(np.float64)
    for s_ in range(planes_last.shape[0]):
        a = planes_last[s_].astype(np.float64)
        sel = seg == s_
        if sel.any() and np.all(np.isfinite(a)):
            out[sel] = X[sel] @ a[:3] - a[3] - float(thr) * np.linalg.norm(a[:3])
    return out


def check_segmentation(pts, labels, ground_idx, obstacle_idx, seg_cfg=None, labels_in=None, strict=False, report=None):
    """Asserts the north-star tolerance and returns the number of labels that differ from the oracle's.

    Rule (BASELINE.json north_star): at most FLIP_FRACTION of the points may differ, and a differing point must lie
    within FLIP_BAND_M of the ORACLE's decision surface (the last plane of its partition, evaluated in float64).
    The oracle itself is float32 with sequential sums (what Eigen's strided colwise().mean() does; its GEMM order is
    unknowable here), so on large partitions its own plane drifts from the exact one by more than the band (merged
    1M-point clouds: d off by 3.5e-4 m, tests/golden/f64_model.py). A differing point outside the band is therefore
    accepted only when `strict` is False AND the oracle provably mis-rounds that very point: the independent float64
    evaluation of the same rule (segment_f64) disagrees with the oracle there and agrees with the device.
    `report` (dict) receives flips, max_gap_m (largest distance of a differing point to the oracle's surface),
    outside_band (how many needed the second clause) and the same two figures against the float64 model."""
    from tests.golden.f64_model import segment_f64

    cfg = seg_cfg or O.default_seg_cfg()
    ref = O.segment(pts, cfg, tie_mode=1, labels_in=labels_in)
    n = pts.shape[0]
    assert labels.shape[0] == n
    diff = np.nonzero(labels != ref["labels"])[0]
    # structural invariants that hold regardless of float rounding
    assert np.array_equal(np.sort(np.concatenate([ground_idx, obstacle_idx])),
                          np.sort(np.concatenate([ref["ground_idx"], ref["obstacle_idx"]])))
    assert np.all(labels[ground_idx] == O.GROUND) and np.all(labels[obstacle_idx] == O.OBSTACLE)
    for idx in (ground_idx, obstacle_idx):  # both clouds are x-ascending (segment by segment)
        assert np.all(np.diff(pts[idx, 0]) >= 0)
    assert diff.size <= max(0, int(FLIP_FRACTION * n)), f"{diff.size} of {n} labels differ"
    info = {"flips": int(diff.size), "max_gap_m": 0.0, "outside_band": 0, "flips_vs_f64": None, "max_gap_f64_m": None}
    if diff.size:
        P = cfg.number_of_planar_partitions
        seg = segment_ids(pts, P)
        gap = np.abs(_margins(pts, ref["planes"][:, -1], seg, cfg.orthogonal_distance_threshold)[diff])
        assert np.all(np.isfinite(gap)), "a label differs in a partition without a fitted plane"
        info["max_gap_m"] = float(gap.max())
        outside = diff[gap >= FLIP_BAND_M]
        info["outside_band"] = int(outside.size)
        m64 = segment_f64(pts, **{name: getattr(cfg, name) for name, _ in cfg._fields_})
        d64 = np.nonzero(labels != m64["labels"])[0]
        info["flips_vs_f64"] = int(d64.size)
        info["max_gap_f64_m"] = float(np.abs(m64["margin"][d64]).max()) if d64.size else 0.0
        if outside.size:
            assert not strict, f"{outside.size} labels differ more than {FLIP_BAND_M} m from the oracle's surface (max {gap.max():.3g} m)"
            assert np.array_equal(labels[outside], m64["labels"][outside]), (
                f"labels differ from the oracle {gap.max():.3g} m from its surface and the float64 model sides with the oracle")
        # against the exact evaluation the device must itself be inside the band, always
        assert d64.size <= max(0, int(FLIP_FRACTION * n))
        assert info["max_gap_f64_m"] < FLIP_BAND_M, f"device differs from the float64 model {info['max_gap_f64_m']:.3g} m from its surface"
    else:
        assert np.array_equal(ground_idx, ref["ground_idx"])
        assert np.array_equal(obstacle_idx, ref["obstacle_idx"])
    if report is not None:
        report.update(info)
    return int(diff.size)


def check_clustering(obs_pts, labels, clu_cfg=None, use_ref=True):
    """Bit-exact: raw labels must equal the oracle's (they are canonical by construction: label k is
    the k-th valid cluster by smallest member index)."""
    cfg = clu_cfg or O.default_clu_cfg()
    want = O.cluster(obs_pts, cfg)
    assert labels.dtype == np.int32 and labels.shape == want.shape
    assert np.array_equal(O.canonicalise(labels), O.canonicalise(want)), "canonical cluster partition differs"
    assert np.array_equal(labels, want), "raw cluster labels differ"
    if use_ref and O.ref_available():
        assert np.array_equal(labels, O.ref_cluster(obs_pts, cfg)), "differs from the unmodified reference Clusterer"
    assert not np.any(labels == O.UNDEFINED)
    return int(labels.max() + 1) if labels.size else 0


# ---- concave outlines: the product's sequential core (csrc/chi_shape.h) compiled for the host -------------------------

_HC = None


def host_checks():
    """tests/host/libhost_checks.so: the host-compilable PRODUCT headers behind a C interface"""
    global _HC
    if _HC is None:
        import ctypes as C

        import __graft_entry__ as ge

        _HC = C.CDLL(str(ge.build_host_checks()))
        _HC.hc_chi_outlines.restype = C.c_longlong
    return _HC


def chi_outlines_host(clusters, sort_mode: int = 0):
    """Concave outlines (clusters of 20 points and more) from csrc/chi_shape.h on the CPU. sort_mode 0 = the device's
    policy, 1 = always the std::sort re-enactment, 2 = always (distance, index). Returns (list of xy[h,2] float32 or
    None where the reference does not deliver / 0-row arrays below 20 points, list of local index arrays,
    (clusters that needed the re-enactment, clusters run))."""
    import ctypes as C

    pts, offsets = O._clusters_csr(clusters)
    k = len(clusters)
    sizes = np.zeros(max(k, 1), np.uint32)
    idx = np.zeros(int(offsets[-1]) + k + 1, np.uint32)
    stats = np.zeros(2, np.uint32)
    n = host_checks().hc_chi_outlines(O._p(pts, C.c_float), O._p(offsets, C.c_uint32), C.c_uint32(k), C.c_uint32(4),
                                      C.c_int(sort_mode), O._p(sizes, C.c_uint32), O._p(idx, C.c_uint32),
                                      C.c_longlong(idx.size), O._p(stats, C.c_uint32))
    assert n >= 0
    out, loc, at = [], [], 0
    for c in range(k):
        h = int(sizes[c])
        if h == 0xFFFFFFFF:
            out.append(None)
            loc.append(None)
            continue
        li = idx[at:at + h].astype(np.int64)
        out.append(pts[int(offsets[c]) + li][:, :2].copy())
        loc.append(li)
        at += h
    return out, loc, (int(stats[0]), int(stats[1]))
