#!/usr/bin/env python
"""GPU box: where the concave outlines (chi_shape.cuh) spend their time - per-cluster phase cycles and the schedule of
the largest tasks, for a 154-frame batch and for single frames. Needs LIDAR_B200_CHI_STATS=1 (set here)."""
import os
import sys
from pathlib import Path

os.environ["LIDAR_B200_CHI_STATS"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
from bench import load_workload  # noqa: E402

pkg = ge.load_package()
frames, name = load_workload("kitti154")
L = pkg.lib()


def report(ctx, label):
    st, n = ctx.last_chi_stats()
    st = st.astype(np.float64)
    t0 = st[:, 1].min()
    dur_ms = (st[:, 2] - st[:, 1]) / 1e6
    print(f"== {label}: {n} tasks, {len(st)} recorded, span of the recorded tasks {(st[:, 2].max() - t0) / 1e6:.1f} ms")
    print("   largest tasks: points, start ms, duration ms, us/point, Mcycles seed/sort/sweep/erosion, triangles")
    for r, d in list(zip(st, dur_ms))[:12]:
        print(f"   {int(r[0]):6d} {(r[1] - t0) / 1e6:8.2f} {d:8.2f} {1e3 * d / r[0]:6.2f}   "
              f"{r[3] / 1e6:7.2f} {r[4] / 1e6:7.2f} {r[5] / 1e6:7.2f} {r[6] / 1e6:7.2f}  {int(r[7])}")
    for lo, hi in ((20, 64), (64, 256), (256, 1024), (1024, 4096), (4096, 1 << 30)):
        sel = (st[:, 0] >= lo) & (st[:, 0] < hi)
        if sel.any():
            cyc = st[sel, 3:7].sum(0)
            print(f"   n in [{lo}, {hi}): {int(sel.sum())} recorded, points {int(st[sel, 0].sum())}, us/point {1e3 * dur_ms[sel].sum() / st[sel, 0].sum():.2f}, "
                  f"cycle share seed/sort/sweep/erosion {np.round(cyc / cyc.sum(), 2).tolist()}, last end {(st[sel, 2].max() - t0) / 1e6:.1f} ms")


if not os.environ.get("CHI_SKIP_BATCH"):
    ctx = pkg.Context(device=0, max_points=sum((f.shape[0] + 31) & ~31 for f in frames), max_frames=len(frames))
    ctx.batch_stage(frames)
    for rep in range(2):
        ctx.batch_run()
        ctx._check(L.lidar_b200_batch_group_clusters(ctx._h), "group")
        ctx._check(L.lidar_b200_batch_hull_outlines(ctx._h, 2), "hull")
        ctx.sync()
    report(ctx, "154-frame batch")
    ctx.close()
one = pkg.Context(device=0, max_points=max(f.shape[0] for f in frames) + 64, max_frames=1)
for f in [int(x) for x in os.environ.get("CHI_FRAMES", "120,0").split(",")]:
    one.batch_stage([frames[f]])
    for rep in range(2):
        one.batch_run()
        one._check(L.lidar_b200_batch_group_clusters(one._h), "group")
        one._check(L.lidar_b200_batch_hull_outlines(one._h, 2), "hull")
        one.sync()
    report(one, f"frame {f} alone")
one.close()
