#!/bin/bash
# round 2, final evidence part 2 (GPU box, 1 GPU): whole GPU suite on the final code, default bench line + reference arm,
# outline timing (1 / 4 / 8 batches in flight), per-cluster phase statistics of the concave outlines.
set -u
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout -k 10 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo "bench exit: $?"; head -c 400 gpurun_out/bench_r2_final.json; echo; tail -3 gpurun_out/bench_r2_final.err
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_ref.json 2> gpurun_out/bench_r2_final_ref.err
echo "ref exit: $?"; cut -c1-260 gpurun_out/bench_r2_final_ref.json
timeout -k 10 900 python tools/outline_timing.py kitti154 > gpurun_out/outline_timing.json 2> gpurun_out/outline_timing.err
echo "timing exit: $?"; cat gpurun_out/outline_timing.json; tail -3 gpurun_out/outline_timing.err
CHI_FRAMES=120,0 timeout -k 5 300 python tools/chi_stats.py > gpurun_out/chi_stats_final.txt 2>&1; grep -A3 "==" gpurun_out/chi_stats_final.txt | head -20
