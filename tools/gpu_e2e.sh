#!/bin/bash
# GPU box: the whole -m gpu suite, then one default bench line without the CPU baseline (gpurun_out/bench_trim.json).
set -u
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
timeout -k 10 600 python bench.py --no-cpu-baseline > gpurun_out/bench_trim.json 2> gpurun_out/bench_trim.err
echo "bench exit: $?"; tail -c 2500 gpurun_out/bench_trim.json | cut -c1-1800; tail -3 gpurun_out/bench_trim.err
