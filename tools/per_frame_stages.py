#!/usr/bin/env python
"""GPU box: per-frame stage times (one frame per batch) -> gpurun_out/per_frame_stages.json"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge  # noqa: E402
from bench import load_workload  # noqa: E402

pkg = ge.load_package()
frames, name = load_workload()
ctx = pkg.Context(device=0, max_points=140_000, max_frames=1)
ctx.set_profiling(True)
rows = []
for i, f in enumerate(frames):
    ctx.batch_stage([f])
    best = None
    for _ in range(3):
        ctx.batch_run()
        ctx.sync()
        st = ctx.last_stage_ms()
        st["total"] = ctx.last_run_ms()
        if best is None or st["total"] < best["total"]:
            best = st
    best["frame"] = i
    rows.append(best)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
json.dump(rows, open(ROOT / "gpurun_out" / "per_frame_stages.json", "w"))
keys = [k for k in rows[0] if k != "frame"]
import statistics
for k in keys:
    v = [r[k] for r in rows]
    print(f"{k:16s} min {min(v):7.3f} p50 {statistics.median(v):7.3f} max {max(v):7.3f} (frame {v.index(max(v))}) sum {sum(v):8.2f}")
