#!/usr/bin/env python
"""GPU box: run a few frames of the workload as one batch, N times (ncu target).
    python tools/one_frame.py <first_frame> <n_frames> <repeats>"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge  # noqa: E402
from bench import load_workload  # noqa: E402

first, n, reps = (int(a) for a in sys.argv[1:4])
pkg = ge.load_package()
frames, _ = load_workload()
sel = frames[first:first + n]
ctx = pkg.Context(device=0, max_points=140_000 * n, max_frames=n)
ctx.set_profiling(True)
ctx.batch_stage(sel)
for _ in range(reps):
    ctx.batch_run()
    ctx.sync()
    print({k: round(v, 3) for k, v in ctx.last_stage_ms().items()}, round(ctx.last_run_ms(), 3))
import os
if os.environ.get("LIDAR_B200_REPLAY_STATS"):
    st = ctx.last_replay_stats()
    order = st[:, 2].argsort()[::-1]
    print("jobs", len(st), "frame members kcycles rounds direct taken seeds cands")
    for r in st[order][:12]:
        print(list(map(int, r)))
