#!/bin/bash
# compute-sanitizer over the lock-free union-find and both replay kernels (GPU box), VERDICT r1 item 1(d):
#   racecheck (shared-memory hazards of the CTA replay rounds), memcheck, synccheck
# over the clustering parity tests. Summaries land in gpurun_out/sanitize_cluster_<tool>.txt.
set -u
mkdir -p gpurun_out
K=${K:-"test_cluster_golden_frames or test_cluster_cta_path_extremes or test_cluster_dense_stress"}
for TOOL in ${TOOLS:-memcheck racecheck synccheck}; do
  timeout -k 10 ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file gpurun_out/sanitize_cluster_$TOOL.log \
     python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 1400 -k "$K" > gpurun_out/pytest_sanitize_cluster_$TOOL.log 2>&1
  rc=$?
  { echo "compute-sanitizer --tool $TOOL over: pytest -k \"$K\""; echo "exit code: $rc";
    tail -3 gpurun_out/pytest_sanitize_cluster_$TOOL.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error" gpurun_out/sanitize_cluster_$TOOL.log | sort | uniq -c | head -20; } > gpurun_out/sanitize_cluster_$TOOL.txt
  cat gpurun_out/sanitize_cluster_$TOOL.txt
done
