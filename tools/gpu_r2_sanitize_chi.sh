#!/bin/bash
# GPU box: compute-sanitizer over the concave-outline kernels on the final code - golden frames (mode 2) and the seeded
# sweep whose lattices go through the std::sort re-enactment (shared-memory windows, staging buffer lock).
set -u
mkdir -p gpurun_out
for TOOL in memcheck racecheck; do
  timeout -k 10 1200 compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file gpurun_out/chi_$TOOL.log \
     python -m pytest tests -m gpu -q --timeout 1100 -k "(outlines_golden_frames and 2) or concave_outlines_sweep" > gpurun_out/pytest_chi_$TOOL.log 2>&1
  echo "$TOOL exit: $?"; tail -2 gpurun_out/pytest_chi_$TOOL.log; tail -3 gpurun_out/chi_$TOOL.log
done
