#!/bin/bash
# N-GPU bench lines on one box: bash tools/gpu_scale.sh "2 4 8"
set -u
mkdir -p gpurun_out
for n in ${1:-8}; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  echo "bench $n gpus exit: $?"; tail -c 1500 gpurun_out/bench_${n}gpu.json; tail -3 gpurun_out/bench_${n}gpu.err
done
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; nproc >> gpurun_out/topo.txt
