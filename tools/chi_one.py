#!/usr/bin/env python
"""GPU box: the concave outlines of ONE frame (default 120: a 22 426-point cluster) a few times - the target of an ncu
capture of chi_outline_kernel."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge  # noqa: E402
from bench import load_workload  # noqa: E402

pkg = ge.load_package()
frames, _ = load_workload("kitti154")
f = int(sys.argv[1]) if len(sys.argv) > 1 else 120
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = pkg.lib()
one = pkg.Context(device=0, max_points=frames[f].shape[0] + 64, max_frames=1)
one.batch_stage([frames[f]])
for _ in range(reps):
    one.batch_run()
    one._check(L.lidar_b200_batch_group_clusters(one._h), "group")
    one._check(L.lidar_b200_batch_hull_outlines(one._h, 2), "hull")
    one.sync()
one.close()
