#!/bin/bash
# round 2, final evidence part 1 (GPU box, 1 GPU): whole GPU suite, compute-sanitizer over the concave-outline kernels,
# default bench line + reference arm.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for TOOL in memcheck racecheck; do
  timeout -k 10 900 compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file gpurun_out/chi_$TOOL.log \
     python -m pytest tests -m gpu -q --timeout 800 -k "outlines_golden_frames and 2" > gpurun_out/pytest_chi_$TOOL.log 2>&1
  echo "$TOOL exit: $?"; tail -2 gpurun_out/pytest_chi_$TOOL.log; tail -4 gpurun_out/chi_$TOOL.log
done
timeout -k 10 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo "bench exit: $?"; tail -c 1500 gpurun_out/bench_r2_final.json; tail -3 gpurun_out/bench_r2_final.err
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_ref.json 2> gpurun_out/bench_r2_final_ref.err
echo "ref exit: $?"; cut -c1-300 gpurun_out/bench_r2_final_ref.json
