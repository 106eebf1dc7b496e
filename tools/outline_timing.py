#!/usr/bin/env python
"""GPU box: time of the device-side cluster split + convex outlines over the bench workload, beside the
reference's own host functions (oracle/_ref, one thread) on a sample of the same clusters."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402
import oracle as O  # noqa: E402  (CPU baseline leg only)
from bench import load_workload  # noqa: E402

pkg = ge.load_package()
which = sys.argv[1] if len(sys.argv) > 1 else "kitti154"
frames, name = load_workload(which)
ctx = pkg.Context(device=0, max_points=sum((f.shape[0] + 31) & ~31 for f in frames), max_frames=len(frames))
ctx.batch_stage(frames)
L = pkg.lib()
out = {"workload": name, "frames": len(frames)}


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    ctx.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return 1e3 * (time.perf_counter() - t0) / reps


def run_only():
    ctx.batch_run()


def run_group():
    ctx.batch_run()
    ctx._check(L.lidar_b200_batch_group_clusters(ctx._h), "group")


def run_group_hull(mode):
    def f():
        run_group()
        ctx._check(L.lidar_b200_batch_hull_outlines(ctx._h, mode), "hull")
    return f


base = timed(run_only)
grp = timed(run_group)
out["ms_per_step"] = {"seg+cluster": base, "+split": grp}
for mode, label in ((0, "+split+convex_outlines"), (1, "+split+concave_small_outlines"), (2, "+split+concave_outlines")):
    out["ms_per_step"][label] = timed(run_group_hull(mode), reps=3 if mode == 2 else 5)
# one frame in flight: the outline step of a single frame (its largest cluster is the tail)
one = pkg.Context(device=0, max_points=max(f.shape[0] for f in frames) + 64, max_frames=1)
lat = []
for f in list(range(0, len(frames), max(1, len(frames) // 12)))[:12]:
    one.batch_stage([frames[f]])
    one.batch_run()
    one._check(L.lidar_b200_batch_group_clusters(one._h), "group")
    one.sync()
    per = []
    for _ in range(3):
        t0 = time.perf_counter()
        one._check(L.lidar_b200_batch_hull_outlines(one._h, 2), "hull")
        one.sync()
        per.append(1e3 * (time.perf_counter() - t0))
    lat.append(min(per))
one.close()
out["concave_outlines_single_frame_ms"] = {"p50": float(np.median(lat)), "max": float(max(lat)), "frames": len(lat)}
# several batches in flight (one context + stream each): the outline step of a batch ends with a few warps on its largest
# clusters, which leaves the GPU to the next batches - the steady-state cost per batch is what a pipeline pays
import threading  # noqa: E402

KMAX = 8
others = [pkg.Context(device=0, max_points=sum((f.shape[0] + 31) & ~31 for f in frames), max_frames=len(frames)) for _ in range(KMAX - 1)]
for o in others:
    o.batch_stage(frames)
ctxs = [ctx] + others


def loop(c, reps):
    for _ in range(reps):
        c.batch_run()
        c._check(L.lidar_b200_batch_group_clusters(c._h), "group")
        c._check(L.lidar_b200_batch_hull_outlines(c._h, 2), "hull")
        c.sync()


for c in ctxs:
    loop(c, 1)
reps = 4
for K in (4, 8):
    th = [threading.Thread(target=loop, args=(c, reps)) for c in ctxs[:K]]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    out["ms_per_step"][f"+split+concave_outlines, {K} batches in flight"] = 1e3 * (time.perf_counter() - t0) / (reps * K)
for o in others:
    o.close()
# results once, for the counts and the CPU sample
res = ctx.batch_fetch()
groups = ctx.batch_clusters()
hulls = ctx.batch_hulls(0)
out["clusters"] = int(sum(g["n_clusters"] for g in groups))
out["outline_vertices"] = int(sum(h["xy"].shape[0] for h in hulls))
out["largest_cluster"] = int(max((np.diff(g["offsets"].astype(np.int64)).max() if g["n_clusters"] else 0) for g in groups))
sample = list(range(0, len(frames), max(1, len(frames) // 8)))[:8]
t_cpu = 0.0
n_cl = 0
for f in sample:
    go = groups[f]["offsets"].astype(np.int64)
    cl = [groups[f]["points"][go[c]:go[c + 1], :3] for c in range(groups[f]["n_clusters"])]
    if O.ref_hull_available():
        t0 = time.perf_counter()
        O.ref_outlines(cl, 0)
        t_cpu += time.perf_counter() - t0
        n_cl += len(cl)
t_cc = 0.0
for f in sample:
    go = groups[f]["offsets"].astype(np.int64)
    cl = [groups[f]["points"][go[c]:go[c + 1], :3] for c in range(groups[f]["n_clusters"])]
    if O.ref_hull_available():
        t0 = time.perf_counter()
        O.ref_outlines(cl, 1)
        t_cc += time.perf_counter() - t0
out["cpu_reference_concave_outlines_ms_per_frame"] = 1e3 * t_cc / max(len(sample), 1)
out["cpu_reference_convex_outlines_ms_per_frame"] = 1e3 * t_cpu / max(len(sample), 1)
out["cpu_sample"] = f"{len(sample)} frames, {n_cl} clusters, 1 thread, findOrderedConvexOutlines via oracle/_ref (includes the ctypes marshalling of the clusters)"
print(json.dumps(out))
