#!/usr/bin/env python
"""Benchmark of the hot path: ground segmentation + Fast Euclidean Clustering.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path over the whole workload (BASELINE.json configs[1]: the
reference's 154-frame data/*.pcd sequence, ~121.7k points per frame; when the frame cache is not on
the box a synthetic 64-beam sequence of the same shape — BASELINE.json configs[4] generator — is
used and named in config.workload). Metric: frames/s (whole job, all GPUs).

  value         device-resident: frames already in HBM, K passes of all kernels, timed with CUDA
                events recorded on the library's own stream (max over ranks)
  e2e           same metric through the public host API (host buffers -> pinned staging -> H2D ->
                kernels -> D2H -> host arrays every step), wall clock around the calls
  roofline      dominant stage's kernel time vs the algorithmic bytes of SURVEY.md §8(d)
  cpu_baseline  the reference's CPU path on this box's host cores (rank 0, N=1), all 154 frames, warmed
  workloads     (N=1) compact sub-lines for the synthetic shapes of SURVEY 8(d) configs 3, 4, 5:
                value, e2e, roofline fraction each (also copied into roofline.other_workloads)
  config5       BASELINE.json configs[4] as specified: a 4096-frame job, rank g of N takes frames
                [g*4096/N, (g+1)*4096/N) - STRONG scaling, reported beside the weak-scaling metric
                (also copied into e2e.config5_strong)
  --impl reference   times only that CPU path (restated Segmenter + UNMODIFIED reference Clusterer
                from oracle/_ref; this is the one place besides tests/smoke that executes oracle/)

Multi-GPU: one process per GPU (torchrun), frames are independent so every rank processes its own
copy of the per-GPU workload with no collective on the data path ("weak" scaling); NCCL is used only
for the barrier and the max-over-ranks of the timings.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "frames/s seg+cluster (154-frame HDL-64E sequence, ~121.7k pts/frame)"
HBM_FALLBACK_GBS = 6650.0


WORKLOADS = ("kitti154", "synth128", "merged1m", "synth64")


def load_workload(which: str = "kitti154"):
    """kitti154 = BASELINE.json configs[1], the configuration the metric is quoted on (default). The others
    are the synthetic clouds of the named shapes, reported as extra lines."""
    from tests.synth import make_frame, make_frame_128, make_merged_1m

    if which == "synth128":
        frames = [make_frame_128(12345 + i) for i in range(64)]
        return frames, "synth128: config 3, 64 synthetic 128-beam frames (128 x 2048 rays, seed 12345 + i), one batch"
    if which == "merged1m":
        frames = [make_merged_1m(777 + i) for i in range(16)]
        return frames, "merged1m: config 4, 16 merged 4-sensor clouds of ~1.04 M points with blobs and lattice walls (seed 777 + i)"
    if which == "synth64":
        distinct = [make_frame(1000 + i) for i in range(64)]
        frames = [distinct[i % 64] for i in range(256)]
        return frames, ("synth64: config 5, synthetic 64-beam frames (seed 1000 + i mod 64) in batches of 256 per GPU; "
                        "the 4096-frame job is 16 such steps split over the GPUs")
    cache = ROOT / "data_cache" / "frames_mm.xz"
    if cache.exists():
        from tools.pack_reference_frames import unpack

        frames = unpack(cache)
        name = "kitti154: reference data/*.pcd sequence (154 frames, 98.5k-124.1k pts, bit-lossless cache)"
    else:
        distinct = [make_frame(1000 + i) for i in range(22)]
        frames = [distinct[i % len(distinct)] for i in range(154)]
        name = "synth64x154: 154 synthetic 64-beam frames (22 distinct scenes cycled), reference data cache absent"
    return frames, name


def check_fingerprints(results, frame_ids, workload):
    """Bit-exact check of the timed run's own outputs against the committed per-frame fingerprints
    (tests/golden/fingerprints.json: oracle segmentation + UNMODIFIED reference Clusterer, made in the
    build container by tests/golden/make_golden.py). numpy only - no oracle code runs here."""
    fp = ROOT / "tests" / "golden" / "fingerprints.json"
    if not workload.startswith("kitti") or not fp.exists():
        return None
    from tools.checksums import mix64

    rows = json.loads(fp.read_text())["frames"]
    if "cluster_labels_mix64" not in rows[0]:
        return None
    out = {"frames_checked": 0, "seg_labels_equal": 0, "same_obstacle_cloud": 0, "cluster_labels_equal_on_same_obstacle_cloud": 0}
    for r, fid in zip(results, frame_ids):
        row = rows[fid % len(rows)]
        out["frames_checked"] += 1
        out["seg_labels_equal"] += int(mix64(r["seg_labels"]) == row["seg_labels_mix64"])
        same = mix64(r["obstacle_idx"]) == row["obstacle_idx_mix64"]
        out["same_obstacle_cloud"] += int(same)
        out["cluster_labels_equal_on_same_obstacle_cloud"] += int(same and mix64(r["cluster_labels"]) == row["cluster_labels_mix64"])
    out["what"] = ("outputs of the timed resident run vs committed fingerprints (oracle segmentation + unmodified reference "
                   "Clusterer, raw labels, bit-exact). Frames whose ground mask differs from the oracle's by a few points "
                   "inside the stated 1e-4 m band feed a different obstacle cloud to the clusterer; those are checked "
                   "against the reference on their own cloud by tests/test_gpu_parity.py::test_all_154_reference_frames")
    return out


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions: an in-process NVML thread
    (nvidia_ml_py, one sample every ~4 ms); `nvidia-smi -lms` is the fallback when NVML cannot load."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index = index
        self.samples = []  # (perf_counter, sm_mhz, reasons bitmask)
        self.windows = []  # timed regions [(t0, t1)]
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self.source = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop.is_set():
                    try:
                        mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                        try:
                            why = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                        except Exception:
                            why = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.samples.append((time.perf_counter(), mhz, why))
                    except Exception:
                        pass
                    time.sleep(0.004)

            self.source = "nvml"
            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            return
        except Exception:
            pass
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                     "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for line in proc.stdout:
                    parts = [x.strip() for x in line.split(",")]
                    try:
                        mhz, self.sm_max = float(parts[0]), float(parts[1])
                    except (ValueError, IndexError):
                        continue
                    why = 0
                    for (_, bit), val in zip(self.REASONS, parts[2:6]):
                        if val.lower().startswith("active"):
                            why |= bit
                    self.samples.append((time.perf_counter(), mhz, why))
                    if self._stop.is_set():
                        break
                proc.terminate()

            self.source = "nvidia-smi"
            self._thread = threading.Thread(target=pump, daemon=True)
            self._thread.start()
            t_end = time.perf_counter() + 5.0
            while not self.samples and time.perf_counter() < t_end:
                time.sleep(0.05)  # nvidia-smi needs a moment before its first line
        except Exception:
            self.source = None

    def window(self, t0: float, t1: float):
        self.windows.append((t0, t1))

    def stop(self):
        self._stop.set()
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)]
        used = inside or self.samples
        if not used:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        why = 0
        for s_ in used:
            why |= s_[2]
        return {"sm_mhz": statistics.median(s_[1] for s_ in used), "sm_max_mhz": self.sm_max,
                "reasons": [name for name, bit in self.REASONS if why & bit], "samples": len(used),
                "samples_inside_timed_regions": len(inside), "source": self.source}


def config_of(workload, nf, total_pts, padded, world):
    """the `config` object: identical keys and values in both arms (the driver compares them)"""
    return {"workload": workload, "frames_per_step_per_gpu": nf, "points_per_step_per_gpu": total_pts,
            "l2": "inputs larger than L2 (%.0f MB of points per step)" % (padded * 16 / 1e6),
            "parallelism": f"frame-sharded x{world}, no collective"}


def run_reference(args, frames, workload):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, every step = ALL frames of
    the workload (the same step as the CUDA arm's), long-lived worker threads and Clusterers (nothing constructed inside
    a timed step)."""
    import oracle as O

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    kind = "reference" if O.ref_available() else "port"
    res = O.ref_pipeline_passes(frames, threads, args.warmup + args.steps)
    total = float(res["pass_wall_s"][args.warmup:].sum())
    nf = len(frames)
    fps = nf * args.steps / total
    pts = int(sum(f.shape[0] for f in frames))
    padded = int(sum((f.shape[0] + 31) & ~31 for f in frames))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "real" if workload.startswith("kitti") else "synthetic",
        "config": config_of(workload, nf, pts, padded, args.gpus),
        "mpts_per_s": pts * args.steps / total / 1e6,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"all {nf} frames per step on {threads} long-lived std::threads: restated Segmenter "
                                   "(Eigen/PCL absent) + unmodified reference Clusterer (oracle/_ref); TBB absent so "
                                   "the reference's par sorts run serially"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


class Env:
    """rank plumbing: barrier + max-over-ranks (NCCL is used for nothing else)"""

    def __init__(self, torch):
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        from lidar_processing_b200 import sharding  # the helper the world-size-2 gloo test covers

        return sharding.reduce_max(x, self.dist, device="cuda")

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def pin_to_gpu_numa_node(local_rank: int):
    """One feeder process per GPU, kept on the host cores next to it: page-locked buffers allocated
    afterwards come from that node. A no-op when the box does not expose the topology (VMs often report -1)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(Path(f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node").read_text())
        if node < 0:
            return "numa node not exposed"
        cpus = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, ids)
        return f"node {node} ({len(ids)} cpus)"
    except Exception as e:  # topology is an optimisation, never a requirement
        return f"unavailable ({type(e).__name__})"


def measure_resident(pkg, env, frames, steps, warmup, sampler):
    """device-resident throughput: frames staged once, `steps` passes of all kernels, CUDA events on the library's stream"""
    nf = len(frames)
    padded = int(sum((f.shape[0] + 31) & ~31 for f in frames))
    ctx = pkg.Context(device=env.local_rank, max_points=padded, max_frames=nf)
    ctx.set_profiling(True)  # (ten event records per step)
    ctx.batch_stage(frames)  # inputs resident in HBM before the timed region
    for _ in range(warmup):
        ctx.batch_run()
    ctx.sync()
    launches0 = ctx.launch_count()
    env.barrier()
    # the K steps are enqueued back to back and timed by two CUDA events on the library's stream around all of them. (A
    # host round trip after every step - synchronise, read the step's events, enqueue the next ~54 launches over four
    # streams just in time - made every step 0.9 ms longer ON THE DEVICE: 14.55 against 13.65 ms per 154 frames.)
    t0 = time.perf_counter()
    ctx.region_begin()
    for _ in range(steps):
        ctx.batch_run()
    region_ms = ctx.region_end_ms()
    env.barrier()
    wall = time.perf_counter() - t0
    if sampler is not None:
        sampler.window(t0, t0 + wall)
    launches = ctx.launch_count() - launches0
    # per-stage times: a profiled pass right behind the timed region, one synchronisation per step so that the stage
    # events of every step can be read (the stage events of the timed steps overwrite each other)
    prof_steps = max(1, min(steps, 5))
    stage_acc = {}
    for _ in range(prof_steps):
        ctx.batch_run()
        ctx.sync()
        for k, v in ctx.last_stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    res = ctx.batch_fetch()
    dev_ms_total = env.max_over_ranks(region_ms)
    return dict(ctx=ctx, res=res, dev_ms_total=dev_ms_total, launches=launches, wall=wall,
                stage_ms={k: v / prof_steps for k, v in stage_acc.items()},
                n_obstacle=int(sum(r["obstacle_idx"].size for r in res)), n_clusters=int(sum(r["n_clusters"] for r in res)))


def measure_e2e(pkg, env, pipe, src, steps, warmup, sampler):
    """end to end through lidar_b200_pipe_*: host clouds -> (H2D) -> kernels -> (D2H) -> host result arrays every step;
    steps are submitted back to back (results of step s go to arena s % 2 and stay readable while step s+1 runs); the
    timed region ends when the last step's results are in host memory"""
    nf = len(src)
    chunks_per_submit = max(1, -(-nf // pipe.chunk_frames))
    n_warm = max(2, min(warmup, 3), -(-pipe.depth // chunks_per_submit))
    for w_ in range(n_warm):
        pipe.submit(src, arena=w_ % 2)  # both result arenas and all `depth` contexts exist before the timed region
    pipe.drain()
    env.barrier()
    l0 = pipe.launch_count()
    t0 = time.perf_counter()
    for s_ in range(steps):
        job = pipe.submit(src, arena=s_ % 2)
    pipe.drain()
    env.barrier()
    t1 = time.perf_counter()
    if sampler is not None:
        sampler.window(t0, t1)
    return env.max_over_ranks(t1 - t0), pipe.results(job), (pipe.launch_count() - l0) // max(steps, 1)


def roofline_of(stage_ms, algo_bytes, step_ms, nf):
    peak, peak_src = hbm_peak()
    dom = max(stage_ms, key=stage_ms.get) if stage_ms else "n/a"
    dom_ms = stage_ms.get(dom, 0.0)
    achieved = algo_bytes / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:
        t = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(dom)
        if t:
            traffic = t["dram_bytes"] / t["frames_in_capture"] * nf
            traffic_src = (f"{t['kernel']}: {t['dram_bytes']} B DRAM read+write in a {t['frames_in_capture']}-frame ncu capture "
                           f"({t['capture']}), scaled per frame to this launch's {nf} frames")
    except Exception:
        pass
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "kernel": dom, "kernel_ms_per_step": dom_ms, "peak_source": peak_src,
            "algorithmic_bytes_per_step": algo_bytes,
            "whole_path_achieved_GBs": algo_bytes / (step_ms / 1e3) / 1e9 if step_ms > 0 else 0.0,
            "stage_ms_per_step": stage_ms}


def sub_workload(pkg, env, name, args):
    """compact line for one of the synthetic shapes (N=1 only): value, e2e, roofline fraction"""
    frames, workload = load_workload(name)
    steps, warmup = 5, 3
    r = measure_resident(pkg, env, frames, steps, warmup, None)
    nf = len(frames)
    total_pts = int(sum(f.shape[0] for f in frames))
    step_ms = r["dev_ms_total"] / steps
    roof = roofline_of(r["stage_ms"], 20 * total_pts + 4 * r["n_obstacle"], step_ms, nf)
    lat = []
    r["ctx"].set_profiling(False)
    for f in frames[: min(nf, 12)]:
        t1 = time.perf_counter()
        r["ctx"].process_batch([f])
        lat.append(1e3 * (time.perf_counter() - t1))
    r["ctx"].close()
    max_chunk_pts = max(sum((f.shape[0] + 31) & ~31 for f in frames[a:a + args.chunk]) for a in range(0, nf, args.chunk))  # (an upper bound: the pipe cuts equal chunks of at most args.chunk frames)
    pipe = pkg.FramePipeline(device=env.local_rank, depth=args.depth, chunk_frames=args.chunk, max_points_per_chunk=max_chunk_pts)
    e2e_s, _, _ = measure_e2e(pkg, env, pipe, pkg.pin_frames(frames), steps, warmup, None)
    pipe.close()
    return {"workload": workload, "value": nf * steps / (r["dev_ms_total"] / 1e3), "unit": "frames/s",
            "mpts_per_s": total_pts * steps / (r["dev_ms_total"] / 1e3) / 1e6, "ms_per_step": step_ms,
            "e2e": nf * steps / e2e_s, "frac": roof["frac"], "kernel": roof["kernel"], "kernel_ms_per_step": roof["kernel_ms_per_step"],
            "stage_ms_per_step": r["stage_ms"], "latency_ms_p50": statistics.median(lat[2:] or lat), "steps": steps, "warmup": warmup}


def dropin_timing(frames):
    """The REAL drop-in call (reference src/processor.cpp:150-178): C++ Segmenter::segment + Clusterer::cluster on pageable
    pcl::PointCloud<pcl::PointXYZI> (32-byte records), one frame at a time, synchronous - tests/cpp/bench_dropin.cpp."""
    import tempfile

    import __graft_entry__ as ge

    exe = ge.build_dropin_bench()
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "frames.bin"
        with open(path, "wb") as f:
            f.write(np.uint32(len(frames)).tobytes())
            f.write(np.array([fr.shape[0] for fr in frames], np.uint32).tobytes())
            for fr in frames:
                f.write(np.ascontiguousarray(fr, np.float32).tobytes())
        r = subprocess.run([str(exe), str(path), "2"], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        return {"error": (r.stderr or r.stdout)[-300:]}
    return json.loads(r.stdout.strip().splitlines()[-1])


def config5_strong(pkg, env, args):
    """BASELINE.json configs[4] / SURVEY 8(d) config 5 as specified: a job of 4096 frames of the 64-beam generator, rank g of
    N takes frames [g*4096/N, (g+1)*4096/N) (contiguous blocks, sharding.shard_frames) - STRONG scaling: the job is fixed,
    the time is the slowest rank's. 64 distinct scenes (seed 1000 + i mod 64) are generated on the host and cycled;
    resident: the rank's frames are processed in batches of 256 staged in HBM beforehand; e2e: the rank's frames go
    through the frame pipeline from page-locked host clouds into page-locked result arrays."""
    from lidar_processing_b200 import sharding
    from tests.synth import make_frame

    job = 4096
    distinct = [make_frame(1000 + i) for i in range(64)]
    mine = sharding.shard_frames(job, env.rank, env.world)
    batch = min(256, len(mine))
    frames = [distinct[i % 64] for i in mine[:batch]]  # every batch of the rank has the same composition (i mod 64)
    n_batches = -(-len(mine) // batch)
    padded = int(sum((f.shape[0] + 31) & ~31 for f in frames))
    ctx = pkg.Context(device=env.local_rank, max_points=padded, max_frames=batch)
    ctx.batch_stage(frames)
    for _ in range(3):
        ctx.batch_run()
    ctx.sync()
    env.barrier()
    ctx.region_begin()  # the rank's batches back to back, two CUDA events on the library's stream around all of them
    for _ in range(n_batches):
        ctx.batch_run()
    ms = ctx.region_end_ms()
    env.barrier()
    ms = env.max_over_ranks(ms)
    ctx.close()
    pinned = pkg.pin_frames(distinct)
    src = [pinned[i % 64] for i in mine]
    pipe = pkg.FramePipeline(device=env.local_rank, depth=args.depth, chunk_frames=args.chunk, gpus_sharing_host=env.world)
    subs = [src[a:a + batch] for a in range(0, len(src), batch)]  # sub-jobs of one batch each, result arenas alternate
    for w_ in range(2):
        pipe.submit(subs[w_ % len(subs)], arena=w_)  # both result arenas and every context exist before the timed region
    pipe.drain()
    env.barrier()
    t0 = time.perf_counter()
    for k, sub in enumerate(subs):
        pipe.submit(sub, arena=k % 2)
    pipe.drain()
    env.barrier()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    pipe.close()
    pts = int(sum(distinct[i % 64].shape[0] for i in range(job)))
    return {"what": "4096-frame job of synthetic 64-beam frames (seed 1000 + i mod 64), rank g takes [g*4096/N, (g+1)*4096/N)",
            "scaling": "strong", "frames": job, "n_gpus": env.world, "frames_per_gpu": len(mine),
            "value": job / (ms / 1e3), "unit": "frames/s", "job_ms": ms, "mpts_per_s": pts / (ms / 1e3) / 1e6,
            "e2e": job / e2e_s, "e2e_job_ms": 1e3 * e2e_s}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="use only the first N frames of the workload (debug)")
    ap.add_argument("--frame-offset", type=int, default=0, help="skip the first N frames (debug / profiling)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-workload lines, config 5 and the next-row timings")
    # (77 x 4 since the window-synchronous replay: larger batches amortise its per-launch tail; 22 x 6 before. Measured
    # at 1 GPU, 154-frame job: 22x6 9 641, 31x5 10 114, 39x4 10 188, 77x3 10 166-10 526, 77x4 10 675, 154x2 10 275 frames/s,
    # profiles/r02_e2e_chunk_depth_sweep.txt)
    ap.add_argument("--chunk", type=int, default=77, help="frames per pipeline chunk (e2e path)")
    ap.add_argument("--depth", type=int, default=4, help="pipeline depth = contexts in rotation (e2e path)")
    ap.add_argument("--workload", default="kitti154", choices=WORKLOADS,
                    help="kitti154 = the metric's configuration; the rest are the synthetic shapes of SURVEY 8(d)")
    args = ap.parse_args()

    frames, workload = load_workload(args.workload)
    global METRIC
    if args.workload != "kitti154":
        METRIC = f"frames/s seg+cluster ({args.workload}, SURVEY 8(d) synthetic shape)"
    if args.frame_offset or args.frames:
        frames = frames[args.frame_offset:][: (args.frames or None)]
    if args.impl == "reference":
        run_reference(args, frames, workload)
        return

    import torch

    import __graft_entry__ as ge

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    env = Env(torch)
    rank, local_rank, world = env.rank, env.local_rank, env.world
    numa = pin_to_gpu_numa_node(local_rank) if world > 1 else "single process, not pinned"
    pkg = ge.load_package()
    from lidar_processing_b200 import sharding

    # the job is world x (workload) frames, split into contiguous blocks, one per rank; frame id i
    # uses workload frame i mod len(workload), so every rank holds the same amount of work (weak scaling)
    mine = sharding.shard_frames(world * len(frames), rank, world)
    frame_ids = [args.frame_offset + (i % len(frames)) for i in mine]  # index into the 154-frame sequence
    frames = [frames[i % len(frames)] for i in mine]
    nf = len(frames)
    total_pts = int(sum(f.shape[0] for f in frames))
    padded = int(sum((f.shape[0] + 31) & ~31 for f in frames))

    # ---- device-resident throughput (`value`) -------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    r = measure_resident(pkg, env, frames, args.steps, args.warmup, sampler)
    ctx, res, dev_ms_total, launches = r["ctx"], r["res"], r["dev_ms_total"], r["launches"]
    n_obstacle, n_clusters = r["n_obstacle"], r["n_clusters"]
    parity = check_fingerprints(res, frame_ids, workload) if rank == 0 else None
    value = sharding.job_throughput([nf * args.steps] * world, [dev_ms_total / 1e3] * world)  # all frames / slowest rank

    # ---- SURVEY 8(f) rows 1 and 3 on the same resident batch (reported beside the metric, not part of it):
    # extra time per step of the device-side per-cluster split and of the ordered convex outlines behind it
    next_rows = None
    if rank == 0 and not args.no_extras:
        try:
            L = pkg.lib()

            def timed_ms(extra):
                for _ in range(2):
                    ctx.batch_run()
                    extra()
                ctx.sync()
                t_ = time.perf_counter()
                for _ in range(3):
                    ctx.batch_run()
                    extra()
                ctx.sync()
                return 1e3 * (time.perf_counter() - t_) / 3

            def split():
                ctx._check(L.lidar_b200_batch_group_clusters(ctx._h), "batch_group_clusters")

            def split_outlines():
                split()
                ctx._check(L.lidar_b200_batch_hull_outlines(ctx._h, pkg.HULL_CONVEX), "batch_hull_outlines")

            def split_concave_outlines():
                split()
                ctx._check(L.lidar_b200_batch_hull_outlines(ctx._h, pkg.HULL_CONCAVE), "batch_hull_outlines")

            base_ms = timed_ms(lambda: None)
            split_ms = timed_ms(split)
            outl_ms = timed_ms(split_outlines)
            hulls = ctx.batch_hulls(pkg.HULL_CONVEX)
            concave_ms = timed_ms(split_concave_outlines)
            chulls = ctx.batch_hulls(pkg.HULL_CONCAVE, tolerate_open_marches=True)
            next_rows = {"what": "wall-clock ms per resident step, 3 steps each: seg+cluster alone, + device-side cluster split "
                                 "(processor.cpp:180-200), + ordered convex outlines of every cluster (polygon_simplification.cpp:31-79), "
                                 "+ the concave policy instead (polygon_simplification.cpp:81-149: Delaunay-based chi-shape from 20 points on, "
                                 "one batch in flight - its step ends with single warps on the largest clusters)",
                         "seg_cluster_ms": base_ms, "plus_split_ms": split_ms, "plus_split_outlines_ms": outl_ms,
                         "outline_vertices_per_step": int(sum(h["xy"].shape[0] for h in hulls)),
                         "plus_split_concave_outlines_ms": concave_ms,
                         "concave_outline_vertices_per_step": int(sum(h["xy"].shape[0] for h in chulls))}
        except Exception as e:  # the metric must not depend on the extra rows
            next_rows = {"error": str(e)}

    # ---- end to end through the host API ------------------------------------------------------
    # Headline: clouds and result arrays in page-locked host memory (lidar_b200_host_alloc), as a loader feeding the
    # library would keep them; the pageable variant (library stages through its own pinned buffers) is reported too.
    pipe = pkg.FramePipeline(device=local_rank, depth=args.depth, chunk_frames=args.chunk, gpus_sharing_host=world)
    pinned_frames = pkg.pin_frames(frames)
    e2e_s, e2e_out, pipe_launches = measure_e2e(pkg, env, pipe, pinned_frames, args.steps, args.warmup, sampler)
    e2e_fps = world * nf * args.steps / e2e_s
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    fetch_mode = pipe.fetch_mode()
    e2e_same = all(np.array_equal(a["cluster_labels"], b["cluster_labels"]) and np.array_equal(a["seg_labels"], b["seg_labels"])
                   for a, b in zip(e2e_out, res))
    e2e_pageable_s, _, _ = measure_e2e(pkg, env, pipe, frames, args.steps, args.warmup, sampler)
    e2e_pageable_fps = world * nf * args.steps / e2e_pageable_s
    clocks = sampler.stop()
    pipe.close()

    # ---- p50 per-frame latency, one frame in flight (submit -> labels on host), rank 0 at every N ----
    lat = []
    ctx.set_profiling(False)  # (stage events off: a single frame is then replayed as one CUDA-graph launch)
    if rank == 0:
        for f in frames[: min(nf, 48)]:
            t1 = time.perf_counter()
            ctx.process_batch([f])
            lat.append(1e3 * (time.perf_counter() - t1))
        lat = lat[4:] if len(lat) > 8 else lat
    latency = {"p50": statistics.median(lat) if lat else None,
               "p95": sorted(lat)[int(0.95 * (len(lat) - 1))] if lat else None, "frames": len(lat),
               "what": "one frame in flight, host buffers in, labels on host out"}
    ctx.close()

    # ---- roofline of the dominant stage ------------------------------------------------------
    algo_bytes = 16 * total_pts + 4 * total_pts + 4 * n_obstacle  # B = 16N + 4N + 4M per frame
    roofline = roofline_of(r["stage_ms"], algo_bytes, dev_ms_total / args.steps, nf)

    # ---- the other configurations under the driver's eye ---------------------------------------
    c5, subs = None, None
    if not args.no_extras and args.workload == "kitti154" and not (args.frames or args.frame_offset):
        try:
            c5 = config5_strong(pkg, env, args)  # every rank takes part (strong scaling over the N GPUs)
        except Exception as e:
            c5 = {"error": f"{type(e).__name__}: {e}"}
        if world == 1:
            subs = {}
            for name in ("synth128", "merged1m", "synth64"):
                try:
                    subs[name] = sub_workload(pkg, env, name, args)
                except Exception as e:
                    subs[name] = {"error": f"{type(e).__name__}: {e}"}

    if rank != 0:
        env.close()
        return

    dropin = None
    if not args.no_extras and args.workload == "kitti154":
        try:
            dropin = dropin_timing(frames)
        except Exception as e:
            dropin = {"error": f"{type(e).__name__}: {e}"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle as O  # CPU baseline leg only

        threads = os.cpu_count() or 1
        r1 = O.ref_pipeline_passes(frames[:12], 1, 2)           # 1-thread latency, second (warm) pass
        rN = O.ref_pipeline_passes(frames, threads, 3)          # all frames; pass 0 is the warm-up
        wallN = float(rN["pass_wall_s"][1:].mean())
        cpu = {"value": nf / wallN, "unit": "frames/s", "cores": threads,
               "kind": "reference" if O.ref_available() else "port",
               "sample": f"all {nf} frames on {threads} long-lived std::threads, mean of 2 passes after 1 warm-up pass "
                         f"(restated Segmenter + unmodified reference Clusterer, TBB absent); 1-thread latency p50 "
                         f"{statistics.median(r1['per_frame_ms']):.1f} ms/frame over {len(r1['per_frame_ms'])} frames",
               "one_thread_ms_per_frame_p50": statistics.median(r1["per_frame_ms"])}

    def compact(d):
        return None if d is None else {k: ({kk: vv for kk, vv in v.items() if kk in ("value", "e2e", "frac", "ms_per_step", "error")}
                                           if isinstance(v, dict) else v) for k, v in d.items()}

    roofline["other_workloads"] = compact(subs)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "real" if workload.startswith("kitti") else "synthetic",
        "config": config_of(workload, nf, total_pts, padded, world),
        "mpts_per_s": world * total_pts * args.steps / (dev_ms_total / 1e3) / 1e6,
        "latency_ms": latency,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": f"lidar_b200_pipe_* : {args.chunk}-frame chunks x depth {args.depth}, clouds and results in "
                        "page-locked host memory, H2D + kernels + D2H every step, wall clock",
                "pageable_host_buffers_value": e2e_pageable_fps, "results_equal_resident_run": bool(e2e_same),
                "gpu_launches_per_step": int(pipe_launches), "latency_ms_p50": latency["p50"], "latency_ms_p95": latency["p95"],
                "host_affinity": numa,
                "fetch_mode": fetch_mode,
                "fetch_mode_what": "0 = copy engines, four result arrays at slot size; 4 = one kernel writes the used part "
                                   "of every slot into the page-locked result arrays (from 3 GPUs sharing the host's DMA "
                                   "path on: lidar_b200_pipe_set_host_sharing(world size)); LIDAR_B200_FETCH_MODE overrides",
                "dropin_value": None if not dropin else dropin.get("frames_per_s"),
                "dropin_p50_ms": None if not dropin else dropin.get("p50_ms"),
                "dropin_what": "C++ Segmenter::segment + Clusterer::cluster (dropin/*.hpp) on pageable 32-byte pcl::PointXYZI "
                               "clouds, one frame at a time, synchronous (tests/cpp/bench_dropin.cpp)",
                "config5_strong": None if c5 is None else {k: c5.get(k) for k in ("value", "e2e", "frames", "n_gpus", "scaling", "error") if k in c5}},
        "gpu_launches": int(launches),
        "next_rows": next_rows,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
        "dropin": dropin,
        "workloads": subs,
        "config5": c5,
        "results": {"obstacle_points_per_step": n_obstacle, "clusters_per_step": n_clusters,
                    "wall_s_resident_region": r["wall"]},
    }
    print(json.dumps(line), flush=True)
    env.close()


if __name__ == "__main__":
    main()
