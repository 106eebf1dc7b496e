#!/usr/bin/env python
"""GPU box: device-resident throughput with the workload split over K contexts running concurrently."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
from bench import load_workload  # noqa: E402

pkg = ge.load_package()
frames, _ = load_workload()
for K in (1, 2, 3, 4, 7):
    parts = [frames[i::K] for i in range(K)]
    ctxs = [pkg.Context(device=0, max_points=sum((f.shape[0] + 31) & ~31 for f in p), max_frames=len(p)) for p in parts]
    for c, p in zip(ctxs, parts):
        c.batch_stage(p)
    for _ in range(3):
        for c in ctxs:
            c.batch_run()
        for c in ctxs:
            c.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 5
    for _ in range(steps):
        for c in ctxs:
            c.batch_run()
    for c in ctxs:
        c.sync()
    dt = (time.perf_counter() - t0) / steps
    print(f"K={K}: {1e3*dt:.2f} ms per 154 frames -> {154/dt:.0f} frames/s")
    for c in ctxs:
        c.close()
