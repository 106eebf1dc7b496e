"""World-size-2 test of the multi-GPU host logic on CPU (gloo): frame sharding covers the job
exactly once with no exchange, and the job time is the max over ranks."""
import os
import socket
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as ge

    ge.load_package()
    from lidar_processing_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_frames = 4096 + 3
    mine = sharding.shard_frames(n_frames, rank, world)
    # every rank learns every shard only to CHECK the cover; the data path itself has no collective
    t = torch.tensor([mine.start, mine.stop], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, t)
    fake_seconds = 1.0 + rank  # rank 1 is the slow one
    slowest = sharding.reduce_max(fake_seconds, dist)
    out.put((rank, [g.tolist() for g in gathered], slowest, len(mine)))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_world2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, shards, slowest, n_mine in results:
        assert shards[0][0] == 0 and shards[0][1] == shards[1][0] and shards[1][1] == 4099  # disjoint, exact cover
        assert slowest == 2.0
        assert n_mine in (2049, 2050)


def test_shard_frames_properties():
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as ge

    ge.load_package()
    from lidar_processing_b200 import sharding

    for n in (0, 1, 7, 154, 4096):
        for world in (1, 2, 4, 8):
            blocks = [sharding.shard_frames(n, r, world) for r in range(world)]
            flat = [i for b in blocks for i in b]
            assert flat == list(range(n))
            assert max(len(b) for b in blocks) - min(len(b) for b in blocks) <= 1
    assert sharding.job_throughput([10, 10], [1.0, 2.0]) == 10.0
    with pytest.raises(ValueError):
        sharding.shard_frames(10, 2, 2)


def test_reference_arm_under_torchrun_prints_one_line():
    """The driver launches `bench.py --impl reference --gpus N` like the GPU arm (torchrun, N ranks): rank 0 alone runs the
    reference's CPU path and prints the line, the other ranks exit 0 without work. CPU only."""
    import json
    import subprocess

    cache = ROOT / "data_cache" / "frames_mm.xz"
    if not cache.exists():
        pytest.skip("data_cache/frames_mm.xz not built (run __graft_entry__.build() with /root/reference present)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = lines[0]
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0
