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
        device-formulation model == ref_cluster.
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

HERE = Path(__file__).resolve().parent


def main():
    paths = O.reference_frame_paths()
    assert len(paths) == 154, "reference data not found"
    pack([paths[0], paths[77], paths[153]], HERE / "frames_0_77_153.xz")
    rows = []
    pinned = dict(oracle_cluster_eq_ref=0, kd_transcription_eq_ref=0, model_eq_ref=0, frames=0)
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
        rows.append(dict(
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
    out = dict(reference_commit="2daa1d1", pinned=pinned, frames=rows)
    (HERE / "fingerprints.json").write_text(json.dumps(out, indent=1))
    print("done", pinned)
    assert pinned["oracle_cluster_eq_ref"] == pinned["kd_transcription_eq_ref"] == pinned["model_eq_ref"] == 154


if __name__ == "__main__":
    main()
