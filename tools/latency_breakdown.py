#!/usr/bin/env python
"""GPU box: where the latency of ONE frame in flight goes - wall clock of process_batch([frame]) against the device time
of its kernels (lidar_b200_last_run_ms, CUDA events around batch_run) and the CUDA-event stage times, over 40 frames.
What is left between the two is uploads, downloads, launch gaps and host work: the most a CUDA Graph / programmatic
dependent launch of the ~60-kernel sequence could remove."""
import json
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge  # noqa: E402
from bench import load_workload  # noqa: E402

pkg = ge.load_package()
frames, _ = load_workload("kitti154")
ctx = pkg.Context(device=0, max_points=max(f.shape[0] for f in frames) + 64, max_frames=1)
for f in frames[:6]:
    ctx.process_batch([f])
wall, dev, launches = [], [], []
for f in frames[6:46]:
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    ctx.process_batch([f])
    wall.append(1e3 * (time.perf_counter() - t0))
    dev.append(ctx.last_run_ms())
    launches.append(ctx.launch_count() - l0)
graph_launches = ctx.graph_launch_count()
ctx.set_profiling(True)
stages = []
for f in frames[6:46]:
    ctx.process_batch([f])
    stages.append(ctx.last_stage_ms())
ctx.close()
out = {"frames": len(wall), "wall_ms_p50": statistics.median(wall), "device_ms_batch_run_p50": statistics.median(dev),
       "launches_per_frame_p50": statistics.median(launches), "frames_replayed_as_one_graph_launch": graph_launches,
       "stage_ms_p50": {k: statistics.median(s[k] for s in stages) for k in stages[0]},
       "outside_the_kernels_ms_p50": statistics.median(w - d for w, d in zip(wall, dev))}
print(json.dumps(out))
