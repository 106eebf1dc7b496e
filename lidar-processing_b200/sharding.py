"""Frame sharding for multi-GPU runs. Frames are independent (the reference keeps no temporal state
on this path), so a job is split into contiguous frame blocks, one block per GPU / process, and no
collective is needed on the data path. torch.distributed is used by callers only
for the barrier and for the max-over-ranks of the timings."""
from __future__ import annotations


def shard_frames(n_frames: int, rank: int, world: int) -> range:
    """Contiguous block of frame ids owned by `rank`: [rank*n/world, (rank+1)*n/world)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    return range(rank * n_frames // world, (rank + 1) * n_frames // world)


def job_throughput(frames_per_rank, seconds_per_rank) -> float:
    """Whole-job frames/s: all frames processed divided by the slowest rank's time."""
    return float(sum(frames_per_rank)) / max(seconds_per_rank)


def reduce_max(value: float, dist=None, device=None) -> float:
    """max over ranks through torch.distributed (any backend); identity without a process group."""
    if dist is None or not dist.is_initialized():
        return value
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
