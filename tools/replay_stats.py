#!/usr/bin/env python
"""Per-job counters of the CTA-per-component replay for the 154-frame batch (GPU box).
usage: LIDAR_B200_REPLAY_STATS=1 python tools/replay_stats.py"""
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
ctx = pkg.Context(device=0, max_points=sum((f.shape[0] + 31) & ~31 for f in frames), max_frames=len(frames))
ctx.set_profiling(True)
ctx.batch_stage(frames)
for _ in range(3):
    ctx.batch_run()
    ctx.sync()
st = ctx.last_replay_stats()
if os.environ.get("LIDAR_B200_REPLAY_V", "5") not in ("1", "2"):  # window-synchronous replay (replay_gen.cuh)
    print("stage ms", {k: round(v, 3) for k, v in ctx.last_stage_ms().items()})
    kc = st[:, 2].astype(np.float64)
    print(f"jobs {st.shape[0]}  sum kcycles {kc.sum():.0f}  max {kc.max():.0f}  mean {kc.mean():.1f}")
    print(f"sum/444 CTAs = {kc.sum()/444:.0f} kcycles = {kc.sum()*1.024/444/1965:.3f} ms ; longest job = {kc.max()*1.024/1965:.3f} ms")
    st = st.copy()
    extra = st[:, 4:6].copy()  # tiles | settle loop, lookups | commit (kcycles, 16 bits each)
    st[:, 4] = st[:, 3] >> 16  # expanded
    st[:, 3] &= 0xFFFF         # windows
    st[:, 5] = 0               # (entries are no longer recorded)
    ph = np.stack([st[:, 6] & 0xFFFF, st[:, 6] >> 16, st[:, 7] & 0xFFFF, st[:, 7] >> 16], 1).astype(np.float64)
    print("top jobs: [frame members kcycles windows expanded entries] kcyc[load settle candidates sort+commit]  cycles/window")
    for j in np.argsort(-kc)[:12]:
        r = st[j]
        print("  ", r[:6].tolist(), ph[j].astype(int).tolist(), round(1024 * r[2] / max(1, r[3])))
    wn = st[:, 3].astype(np.float64)
    print(f"total windows {wn.sum():.0f}, mean cycles/window {1024*kc.sum()/wn.sum():.0f}; entries/window {st[:,5].sum()/wn.sum():.1f} "
          f"expanded/window {st[:,4].sum()/wn.sum():.2f}; phase share load {ph[:,0].sum()/kc.sum():.2f} settle {ph[:,1].sum()/kc.sum():.2f} "
          f"candidates {ph[:,2].sum()/kc.sum():.2f} sort+commit {ph[:,3].sum()/kc.sum():.2f}")
    for lo, hi in ((256, 512), (512, 1024), (1024, 2048), (2048, 4096), (4096, 8192), (8192, 1 << 30)):
        m = (st[:, 1] >= lo) & (st[:, 1] < hi)
        if m.any():
            print(f"members [{lo},{hi}): jobs {int(m.sum())}, kcycles {kc[m].sum():.0f} ({100*kc[m].sum()/kc.sum():.1f} %), cycles/member {1024*kc[m].sum()/st[m,1].sum():.0f}, cycles/window {1024*kc[m].sum()/wn[m].sum():.0f}")
    sys.exit(0)
print("stage ms", {k: round(v, 3) for k, v in ctx.last_stage_ms().items()})
kc = st[:, 2].astype(np.float64)
print(f"jobs {st.shape[0]}  sum kcycles {kc.sum():.0f}  max {kc.max():.0f}  mean {kc.mean():.1f}")
print(f"sum/444 CTAs = {kc.sum()/444:.0f} kcycles = {kc.sum()*1.024/444/1965:.3f} ms ; longest job = {kc.max()*1.024/1965:.3f} ms")
order = np.argsort(-kc)
v2 = os.environ.get("LIDAR_B200_REPLAY_V") == "2"
print("top jobs: [frame members kcycles rounds direct " + ("kcycA kcycB kcycC" if v2 else "kcycA kcycBC kcycEF") + "]  cycles/round   (kcycles = cycles >> 10)")
for j in order[:12]:
    r = st[j]
    print("  ", r.tolist(), round(1024 * r[2] / max(1, r[3] + (0 if v2 else r[4]))))
rounds = st[:, 3].astype(np.float64) + (0 if v2 else st[:, 4])
if v2:
    print(f"total rounds {rounds.sum():.0f}, mean cycles/round {1024*kc.sum()/rounds.sum():.0f}; phase share A (window) "
          f"{st[:,5].sum()/kc.sum():.2f} B (lookup) {st[:,6].sum()/kc.sum():.2f} C (candidates) {st[:,7].sum()/kc.sum():.2f}, rest = write-back")
else:
    print(f"total rounds {rounds.sum():.0f}, mean cycles/round {1024*kc.sum()/rounds.sum():.0f}; phase share A {st[:,5].sum()/kc.sum():.2f} BC {st[:,6].sum()/kc.sum():.2f} EF {st[:,7].sum()/kc.sum():.2f}")
for lo, hi in ((256, 512), (512, 1024), (1024, 2048), (2048, 4096), (4096, 8192), (8192, 1 << 30)):
    m = (st[:, 1] >= lo) & (st[:, 1] < hi)
    if m.any():
        print(f"members [{lo},{hi}): jobs {int(m.sum())}, kcycles {kc[m].sum():.0f} ({100*kc[m].sum()/kc.sum():.1f} %), cycles/member {1024*kc[m].sum()/st[m,1].sum():.0f}")
