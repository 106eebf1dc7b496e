#!/usr/bin/env python
"""Copy-only probe of the e2e byte pattern: no kernels, just the DMA traffic one step of the frame
pipeline moves - 154 H2D copies of one cloud each (16 B / point) and, per 22-frame chunk, the result copies D2H - on two
streams so both copy engines run, one process per GPU (torchrun) exactly like bench.py. Prints frames/s-equivalent and
GB/s per direction, max over ranks: the ceiling the host side of the box puts on `e2e` at this GPU count.

    python tools/copy_probe.py [--d2h-bytes-per-point 16]        (N=1)
    python -m torch.distributed.run --nproc-per-node N ... tools/copy_probe.py
"""
import argparse
import json
import os
import time

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--frames", type=int, default=154)
    ap.add_argument("--points", type=int, default=121733)
    ap.add_argument("--chunk", type=int, default=22)
    ap.add_argument("--d2h-bytes-per-point", type=float, default=16.0, help="16 = four slot-size arrays, 9.7 = exact sizes")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_in = args.points * 16
    n_out = int(args.points * args.d2h_bytes_per_point) * args.chunk
    h_in = [torch.empty(n_in, dtype=torch.uint8).pin_memory() for _ in range(args.frames)]
    d_in = torch.empty(n_in * args.chunk, dtype=torch.uint8, device="cuda")
    n_chunks = -(-args.frames // args.chunk)
    h_out = [torch.empty(n_out, dtype=torch.uint8).pin_memory() for _ in range(n_chunks)]
    d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def step():
        for c in range(n_chunks):
            with torch.cuda.stream(s_in):
                for k, f in enumerate(range(c * args.chunk, min((c + 1) * args.chunk, args.frames))):
                    d_in[k * n_in:(k + 1) * n_in].copy_(h_in[f], non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out[c].copy_(d_out, non_blocking=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        h2d = args.frames * n_in * args.steps
        d2h = n_chunks * n_out * args.steps
        print(json.dumps({"copy_only_frames_per_s": world * args.frames * args.steps / dt, "n_gpus": world,
                          "h2d_GBs_per_gpu": h2d / dt / 1e9, "d2h_GBs_per_gpu": d2h / dt / 1e9,
                          "d2h_bytes_per_point": args.d2h_bytes_per_point, "steps": args.steps}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
