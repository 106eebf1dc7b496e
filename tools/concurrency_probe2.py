#!/usr/bin/env python
"""GPU box: device-resident throughput with K contexts each holding the WHOLE 154-frame batch and stepping at the same
time (the kernels of one batch's tail overlap the next batch's head), against one context stepping alone."""
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
cap = sum((f.shape[0] + 31) & ~31 for f in frames)
for K in (1, 2, 3, 4):
    ctxs = [pkg.Context(device=0, max_points=cap, max_frames=len(frames)) for _ in range(K)]
    for c in ctxs:
        c.batch_stage(frames)
    for _ in range(3):
        for c in ctxs:
            c.batch_run()
    for c in ctxs:
        c.sync()
    torch.cuda.synchronize()
    steps = 10
    t0 = time.perf_counter()
    for _ in range(steps):
        for c in ctxs:
            c.batch_run()
    for c in ctxs:
        c.sync()
    dt = (time.perf_counter() - t0) / (steps * K)
    print(f"K={K}: {1e3*dt:.2f} ms per 154 frames -> {154/dt:.0f} frames/s", flush=True)
    for c in ctxs:
        c.close()
