#!/bin/bash
# round 2, GPU call 4: third-generation CTA replay (live-candidate bitmaps) + fused k-d bottom levels - parity, A/B vs
# the second generation, CTA-threshold sweep, round statistics, per-frame stages.
set -u
mkdir -p gpurun_out
STEPS="tests smoke" PYTEST_ARGS="-x" bash tools/gpu_check.sh
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['roofline']['stage_ms_per_step']
print(sys.argv[2], 'resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'p50', round(d['latency_ms']['p50'],2), 'replay', round(s['replay'],2), 'kd', round(s['kd_order'],2), 'uf', round(s['union_find'],2), 'parity', d['parity']['cluster_labels_equal_on_same_obstacle_cloud'] if d.get('parity') else None)" $1 "$2"; }
for cfg in "3 256 -1" "2 256 -1" "3 128 -1" "3 512 -1"; do
  set -- $cfg
  LIDAR_B200_REPLAY_V=$1 LIDAR_B200_CTA_MIN_MEMBERS=$2 timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sweep4_v$1_m$2_k$3.json 2> gpurun_out/sweep4.err
  summ gpurun_out/sweep4_v$1_m$2_k$3.json "v$1 cta_min $2 kdfused $3:"
done 2>&1 | tee gpurun_out/sweep_replay_r2c4.txt
tail -3 gpurun_out/sweep4.err
LIDAR_B200_REPLAY_STATS=1 timeout 300 python tools/replay_stats.py > gpurun_out/replay_stats_v3.txt 2>&1; cat gpurun_out/replay_stats_v3.txt
timeout 300 python tools/per_frame_stages.py > gpurun_out/per_frame_stages_v3.txt 2>&1; cat gpurun_out/per_frame_stages_v3.txt
