#!/bin/bash
# round 2, GPU call 10: component sort by the frame-resident kernel, the pipe's fetch-mode trials at 1 GPU,
# CTA-threshold sweep for the window-synchronous replay.
set -u
mkdir -p gpurun_out
STEPS="tests" PYTEST_ARGS="-x" PYTEST_TIMEOUT=900 bash tools/gpu_check.sh
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['roofline']['stage_ms_per_step']
print(sys.argv[2], 'resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'fetch', d['e2e'].get('fetch_mode'), 'p50', round(d['latency_ms']['p50'],2), {k: round(v,2) for k,v in s.items()}, 'parity', d['parity']['cluster_labels_equal_on_same_obstacle_cloud'] if d.get('parity') else None)" $1 "$2"; }
for m in 256 128 512 1024; do
  LIDAR_B200_CTA_MIN_MEMBERS=$m timeout -k 10 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sweep10_m$m.json 2> gpurun_out/sweep10.err
  summ gpurun_out/sweep10_m$m.json "cta_min $m:"
done 2>&1 | tee gpurun_out/sweep_r2c10.txt
tail -3 gpurun_out/sweep10.err
