#!/usr/bin/env python
"""GPU box: one 154-frame batch through segmentation + clustering + split + the concave outline policy, twice - the
target of the ncu launch list of the outline kernels."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge  # noqa: E402
from bench import load_workload  # noqa: E402

pkg = ge.load_package()
frames, _ = load_workload("kitti154")
L = pkg.lib()
ctx = pkg.Context(device=0, max_points=sum((f.shape[0] + 31) & ~31 for f in frames), max_frames=len(frames))
ctx.batch_stage(frames)
for _ in range(2):
    ctx.batch_run()
    ctx._check(L.lidar_b200_batch_group_clusters(ctx._h), "group")
    ctx._check(L.lidar_b200_batch_hull_outlines(ctx._h, 2), "hull")
    ctx.sync()
ctx.close()
