#!/bin/bash
# GPU box: phase statistics of the concave outlines, then one ncu --set full capture of chi_outline_kernel on a single
# frame (frame 0: 155 chi-shape clusters, the largest 7372 points).
set -u
mkdir -p gpurun_out
timeout -k 5 300 python tools/chi_stats.py > gpurun_out/chi_stats_v2.txt 2>&1; grep -A3 "==" gpurun_out/chi_stats_v2.txt | head -40
timeout -k 10 500 ncu --set full --clock-control none --import-source on -k regex:chi_outline -s 1 -c 1 \
  -o gpurun_out/ncu_chi_outline_frame0 python tools/chi_one.py 0 2 > gpurun_out/ncu_chi.log 2>&1
echo "ncu exit: $?"; tail -3 gpurun_out/ncu_chi.log
