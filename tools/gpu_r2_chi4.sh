#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "outlines" > gpurun_out/pytest_hull.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_hull.log; tail -4 gpurun_out/pytest_hull.log
timeout -k 5 300 python tools/chi_stats.py > gpurun_out/chi_stats_v4.txt 2>&1; grep -A4 "==" gpurun_out/chi_stats_v4.txt | head -40
timeout -k 10 600 python tools/outline_timing.py kitti154 > gpurun_out/outline_timing.json 2> gpurun_out/outline_timing.err
echo "timing exit: $?"; cat gpurun_out/outline_timing.json; tail -5 gpurun_out/outline_timing.err
