#!/usr/bin/env python
"""Generates the committed golden fixtures from the reference data and the reference build.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py

Outputs (committed):
  tests/golden/frames_0_77_153.xz   three reference frames, lossless (tools/pack_reference_frames.py format)
  tests/golden/fingerprints.json    per-frame counts + FNV-1a and numpy-only mix64 (tools/checksums.py) fingerprints for all 154 frames:
        segmentation  = oracle restatement (stable x order)          [parity unpinned at the Eigen boundary]
        clustering    = UNMODIFIED reference Clusterer (oracle/_ref)  on that obstacle cloud
    and, for every frame, the result of pinning the oracle against the reference build:
        oracle_cluster == ref_cluster, transcribed k-d order == reference k-d order,
        device-formulation model == ref_cluster,
        oracle_segment(tie_mode 0) == the unmodified reference src/segmentation.cpp compiled against the PCL / Eigen
        stand-ins (oracle/eigen_shim: pins everything in that file except Eigen's own floating-point order).
    Every frame row also carries
        planes_f64 / n_ground_f64  the planes of the independent float64 model (tests/golden/f64_model.py:
                                   numpy eigh) per partition x iteration - the pin at the Eigen boundary
        f64_vs_oracle              how the restated float32 oracle compares with that model (label flips, the
                                   largest |margin| of a flipped point, largest plane deviation)
        tie_order                  the reference AS IT COMPILES HERE sorts x with one serial introsort
                                   (src/segmentation.cpp:119, TBB absent): same label set, another order of equal-x
                                   points in the obstacle cloud, hence another input order for the order-dependent
                                   Clusterer. Recorded: positions of the obstacle cloud that differ, clusters on
                                   either side, points whose cluster (named by its smallest ORIGINAL point index)
                                   differs.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import oracle as O  # noqa: E402
from tools.pack_reference_frames import pack  # noqa: E402
from tools.checksums import mix64  # noqa: E402
from tests.golden.f64_model import segment_f64, plane_deviation  # noqa: E402

HERE = Path(__file__).resolve().parent


def clusters_by_original_index(n, obstacle_idx, labels):
    """canonical cluster name per ORIGINAL point: smallest original index of the cluster, -1 invalid, -2 not an obstacle"""
    out = np.full(n, -2, np.int64)
    lab = labels.astype(np.int64)
    valid = lab >= 0
    k = int(lab.max()) + 1 if lab.size else 0
    mins = np.full(max(k, 1), np.iinfo(np.int64).max)
    np.minimum.at(mins, lab[valid], obstacle_idx[valid].astype(np.int64))
    out[obstacle_idx[valid]] = mins[lab[valid]]
    out[obstacle_idx[~valid]] = -1
    return out


def tie_order_row(pts, seg1, lab1):
    seg0 = O.segment(pts, tie_mode=0)
    lab0 = O.ref_cluster(pts[seg0["obstacle_idx"]])
    n = pts.shape[0]
    c0 = clusters_by_original_index(n, seg0["obstacle_idx"], lab0)
    c1 = clusters_by_original_index(n, seg1["obstacle_idx"], lab1)
    return dict(seg_labels_equal=bool(np.array_equal(seg0["labels"], seg1["labels"])),
                obstacle_positions_differ=int((seg0["obstacle_idx"] != seg1["obstacle_idx"]).sum())
                if seg0["obstacle_idx"].size == seg1["obstacle_idx"].size else -1,
                n_clusters_introsort=int(lab0.max() + 1) if lab0.size else 0,
                n_clusters_stable=int(lab1.max() + 1) if lab1.size else 0,
                points_in_other_cluster=int((c0 != c1).sum()))


def f64_row(pts, seg):
    m = segment_f64(pts)
    diff = np.nonzero(m["labels"] != seg["labels"])[0]
    dn = dd = 0.0
    for s_ in range(m["planes"].shape[0]):
        for it in range(m["planes"].shape[1]):
            if np.isfinite(m["planes"][s_, it, 0]) and np.isfinite(seg["planes"][s_, it, 0]):
                a, b = plane_deviation(seg["planes"][s_, it], m["planes"][s_, it])
                dn, dd = max(dn, a), max(dd, b)
    return m, dict(label_flips=int(diff.size), max_abs_margin_of_flips_m=float(np.abs(m["margin"][diff]).max()) if diff.size else 0.0,
                   max_normal_dev=dn, max_d_dev_m=dd)


def segment_equals_reference_source(pts) -> bool:
    """Restated Segmenter (std::sort tie order, tie_mode 0) == the UNMODIFIED reference src/segmentation.cpp compiled
    against the PCL / Eigen stand-ins (oracle/_ref/libref_segment.so): labels and the order of both output clouds."""
    a, b = O.segment(pts, tie_mode=0), O.ref_segment(pts)
    return all(np.array_equal(a[k], b[k]) for k in ("labels", "ground_idx", "obstacle_idx"))


def main():
    paths = O.reference_frame_paths()
    assert len(paths) == 154, "reference data not found"
    pack([paths[0], paths[77], paths[153]], HERE / "frames_0_77_153.xz")
    rows = []
    pinned = dict(oracle_cluster_eq_ref=0, kd_transcription_eq_ref=0, model_eq_ref=0, frames=0, oracle_segment_eq_ref_source=0)
    for i, p in enumerate(paths):
        pts = O.read_pcd(p)
        seg = O.segment(pts, tie_mode=1)
        obs = pts[seg["obstacle_idx"]]
        ref_lab = O.ref_cluster(obs)
        ora_lab = O.cluster(obs)
        ref_order = O.ref_kd_order(obs)
        tx_order = O.kd_order(obs, 1)
        rank = np.empty_like(tx_order)
        rank[tx_order] = np.arange(tx_order.size, dtype=np.uint32)
        model_lab, stats = O.cluster_model(obs, rank)
        pinned["frames"] += 1
        pinned["oracle_cluster_eq_ref"] += int(np.array_equal(ref_lab, ora_lab))
        pinned["kd_transcription_eq_ref"] += int(np.array_equal(ref_order, tx_order))
        pinned["model_eq_ref"] += int(np.array_equal(ref_lab, model_lab))
        pinned["oracle_segment_eq_ref_source"] += int(segment_equals_reference_source(pts))
        m64, f64_info = f64_row(pts, seg)
        rows.append(dict(
            planes_f64=[[[float(v) for v in m64["planes"][s_, it]] for it in range(m64["planes"].shape[1])]
                        for s_ in range(m64["planes"].shape[0])],
            n_ground_f64=m64["n_ground"].tolist(), f64_vs_oracle=f64_info, tie_order=tie_order_row(pts, seg, ref_lab),
            frame=p.name, n=int(pts.shape[0]), n_ground=int(seg["ground_idx"].size),
            n_obstacle=int(seg["obstacle_idx"].size), seg_status=[int(s) for s in seg["status"]],
            seg_labels_fnv=f"{O.fnv1a64(seg['labels']):016x}", obstacle_idx_fnv=f"{O.fnv1a64(seg['obstacle_idx']):016x}",
            planes=[[float(v) for v in seg["planes"][s, -1]] for s in range(seg["planes"].shape[0])],
            n_clusters=int(ref_lab.max() + 1) if ref_lab.size else 0, n_invalid=int((ref_lab == -1).sum()),
            cluster_labels_fnv=f"{O.fnv1a64(ref_lab):016x}", kd_order_fnv=f"{O.fnv1a64(ref_order):016x}",
            replay=stats,
            seg_labels_mix64=mix64(seg["labels"]), obstacle_idx_mix64=mix64(seg["obstacle_idx"]),
            ground_idx_mix64=mix64(seg["ground_idx"]), cluster_labels_mix64=mix64(ref_lab),
        ))
        if i % 10 == 0:
            print(i, rows[-1]["frame"], rows[-1]["n_obstacle"], rows[-1]["n_clusters"], pinned, flush=True)
    tie = dict(frames=len(rows),
               frames_with_equal_label_set=sum(r["tie_order"]["seg_labels_equal"] for r in rows),
               frames_with_other_obstacle_order=sum(r["tie_order"]["obstacle_positions_differ"] != 0 for r in rows),
               frames_with_other_cluster_count=sum(r["tie_order"]["n_clusters_introsort"] != r["tie_order"]["n_clusters_stable"] for r in rows),
               frames_with_other_partition=sum(r["tie_order"]["points_in_other_cluster"] != 0 for r in rows),
               points_in_other_cluster=sum(r["tie_order"]["points_in_other_cluster"] for r in rows),
               obstacle_points=sum(r["n_obstacle"] for r in rows))
    f64 = dict(label_flips=sum(r["f64_vs_oracle"]["label_flips"] for r in rows),
               max_abs_margin_of_flips_m=max(r["f64_vs_oracle"]["max_abs_margin_of_flips_m"] for r in rows),
               max_normal_dev=max(r["f64_vs_oracle"]["max_normal_dev"] for r in rows),
               max_d_dev_m=max(r["f64_vs_oracle"]["max_d_dev_m"] for r in rows), points=sum(r["n"] for r in rows))
    out = dict(reference_commit="2daa1d1", pinned=pinned, tie_order_summary=tie, f64_vs_oracle_summary=f64, frames=rows)
    (HERE / "fingerprints.json").write_text(json.dumps(out, indent=1))
    print("done", pinned)
    assert pinned["oracle_cluster_eq_ref"] == pinned["kd_transcription_eq_ref"] == pinned["model_eq_ref"] == 154


if __name__ == "__main__":
    main()
