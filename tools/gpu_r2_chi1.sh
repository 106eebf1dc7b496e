#!/bin/bash
# GPU box: the outline tests (convex + concave chi-shape), then tools/outline_timing.py on the bench workload.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout -k 10 1200 python -m pytest tests -m gpu -q --timeout 900 -x -k "outlines" > gpurun_out/pytest_hull.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_hull.log; tail -30 gpurun_out/pytest_hull.log
timeout -k 10 600 python tools/outline_timing.py kitti154 > gpurun_out/outline_timing.json 2> gpurun_out/outline_timing.err
echo "timing exit: $?"; cat gpurun_out/outline_timing.json; tail -5 gpurun_out/outline_timing.err
