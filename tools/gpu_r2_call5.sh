#!/bin/bash
# round 2, GPU call 5: fourth-generation replay (warp per component, frame bitmaps shared by the CTA) - parity, A/B
# against generations 2 and 3, CTAs-per-SM sweep, per-frame stages.
set -u
mkdir -p gpurun_out
STEPS="tests smoke" PYTEST_ARGS="-x" bash tools/gpu_check.sh
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['roofline']['stage_ms_per_step']
print(sys.argv[2], 'resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'p50', round(d['latency_ms']['p50'],2), 'replay', round(s['replay'],2), 'kd', round(s['kd_order'],2), 'uf', round(s['union_find'],2), 'sort', round(s['component_sort'],2), 'parity', d['parity']['cluster_labels_equal_on_same_obstacle_cloud'] if d.get('parity') else None)" $1 "$2"; }
for cfg in "4 3" "4 4" "4 2" "2 3" "3 3"; do
  set -- $cfg
  LIDAR_B200_REPLAY_V=$1 LIDAR_B200_REPLAY4_CTAS_PER_SM=$2 timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sweep5_v$1_c$2.json 2> gpurun_out/sweep5.err
  summ gpurun_out/sweep5_v$1_c$2.json "v$1 ctas/sm $2:"
done 2>&1 | tee gpurun_out/sweep_replay_r2c5.txt
tail -3 gpurun_out/sweep5.err
timeout 300 python tools/per_frame_stages.py > gpurun_out/per_frame_stages_v4.txt 2>&1; cat gpurun_out/per_frame_stages_v4.txt
