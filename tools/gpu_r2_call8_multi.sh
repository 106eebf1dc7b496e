#!/bin/bash
# round 2, multi-GPU call (run with gpurun --gpus 8): the copy-only ceiling of the e2e byte pattern at 8 and 4 GPUs
# (slot-size and exact-size D2H), then the bench at 8 / 4 / 2 GPUs with result fetch mode 0 (copies at slot size) and
# mode 4 (one kernel writes the exact sizes into page-locked host memory).
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; nproc >> gpurun_out/topo.txt; free -g >> gpurun_out/topo.txt
run() { n=$1; shift; timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
for n in 8 4; do
  for b in 16 9.7; do
    run $n tools/copy_probe.py --steps 20 --d2h-bytes-per-point $b > gpurun_out/copy_probe_${n}gpu_$b.json 2> gpurun_out/copy_probe.err
    cat gpurun_out/copy_probe_${n}gpu_$b.json; tail -2 gpurun_out/copy_probe.err
  done
done
summ() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); e=d['e2e']
print(sys.argv[2], 'resident', round(d['value']), 'e2e', round(e['value']), 'd2h', e['d2h_bytes_per_step'], 'h2d', e['h2d_bytes_per_step'], 'p50', d.get('latency_ms'))" $1 "$2"; }
for n in 8 4 2; do
  for m in 0 4; do
    LIDAR_B200_FETCH_MODE=$m run $n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_${n}gpu_fetch$m.json 2> gpurun_out/bench_multi.err
    summ gpurun_out/bench_${n}gpu_fetch$m.json "$n gpus fetch mode $m:"; tail -2 gpurun_out/bench_multi.err
  done
done 2>&1 | tee gpurun_out/multi_gpu_r2.txt
