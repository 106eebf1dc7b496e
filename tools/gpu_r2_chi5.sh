#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "outlines" > gpurun_out/pytest_hull.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_hull.log; tail -4 gpurun_out/pytest_hull.log
CHI_FRAMES=145,119,115 timeout -k 5 300 python tools/chi_stats.py > gpurun_out/chi_stats_v5.txt 2>&1; grep -A4 "==" gpurun_out/chi_stats_v5.txt | head -40
