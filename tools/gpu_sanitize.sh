#!/bin/bash
# compute-sanitizer over the outline / packing / split tests (GPU box):  TOOL=memcheck|racecheck  K="<pytest -k expr>"
set -u
mkdir -p gpurun_out
TOOL=${TOOL:-memcheck}
timeout -k 10 1000 compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file gpurun_out/$TOOL.log \
   python -m pytest tests -m gpu -q --timeout 900 -k "${K:-outlines_api_edges or colorized or cluster_split_batch or outlines_golden}" > gpurun_out/pytest_$TOOL.log 2>&1
echo "$TOOL exit: $?"; tail -3 gpurun_out/pytest_$TOOL.log; tail -12 gpurun_out/$TOOL.log
