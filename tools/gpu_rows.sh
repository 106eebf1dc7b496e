#!/bin/bash
# GPU box: the split / outline / packing / drop-in tests only.
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 600 -k "outlines or colorized or dropin or split" > gpurun_out/pytest_rows.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_rows.log; tail -30 gpurun_out/pytest_rows.log
