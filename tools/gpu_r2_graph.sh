#!/bin/bash
# GPU box: the single-frame CUDA-graph path - whole GPU suite with it, latency with and without it
set -u
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
LIDAR_B200_GRAPH=0 timeout 200 python tools/latency_breakdown.py > gpurun_out/latency_graph0.json 2> gpurun_out/latency_graph0.err; cat gpurun_out/latency_graph0.json; tail -2 gpurun_out/latency_graph0.err
LIDAR_B200_GRAPH=1 timeout 200 python tools/latency_breakdown.py > gpurun_out/latency_graph1.json 2> gpurun_out/latency_graph1.err; cat gpurun_out/latency_graph1.json; tail -2 gpurun_out/latency_graph1.err
