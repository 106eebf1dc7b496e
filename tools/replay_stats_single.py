#!/usr/bin/env python
"""Per-job phase counters of the window-synchronous replay with ONE frame in flight (unloaded SMs): the latency
structure of a window.   usage: python tools/replay_stats_single.py [frame ...]"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("LIDAR_B200_REPLAY_STATS", "1")
import __graft_entry__ as ge  # noqa: E402
from bench import load_workload  # noqa: E402

pkg = ge.load_package()
frames, name = load_workload()
ctx = pkg.Context(device=0, max_points=140_000, max_frames=1)
ctx.set_profiling(True)
for fi in [int(a) for a in sys.argv[1:]] or [143, 0]:
    ctx.batch_stage([frames[fi]])
    for _ in range(3):
        ctx.batch_run()
        ctx.sync()
    st = ctx.last_replay_stats()
    ms = ctx.last_stage_ms()
    print(f"frame {fi}: replay stage {ms['replay']:.3f} ms, jobs {st.shape[0]}")
    kc = st[:, 2].astype(np.float64)
    win, exp = st[:, 3] & 0xFFFF, st[:, 3] >> 16
    names = ["load", "settle(all)", "candidates(all)", "sort+commit", "tiles", "settle loop", "lookups", "commit"]
    ph = np.stack([st[:, 6] & 0xFFFF, st[:, 6] >> 16, st[:, 7] & 0xFFFF, st[:, 7] >> 16, st[:, 4] & 0xFFFF, st[:, 4] >> 16,
                   st[:, 5] & 0xFFFF, st[:, 5] >> 16], 1).astype(np.float64)
    print("  [members kcycles windows expanded]  cycles/window:", names)
    for j in np.argsort(-kc)[:6]:
        r = st[j]
        w = max(1, int(win[j]))
        print("  ", [int(r[1]), int(r[2]), int(win[j]), int(exp[j])], [int(1024 * x / w) for x in ph[j]], "total/window", int(1024 * r[2] / w))
