#!/bin/bash
# round 2, GPU call 6: fifth-generation (window-synchronous) replay - parity first, then A/B against generation 2,
# live counters on/off, CTAs per SM, per-job stats, merged1m.
set -u
mkdir -p gpurun_out
STEPS="tests smoke" PYTEST_ARGS="-x" PYTEST_TIMEOUT=900 bash tools/gpu_check.sh
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['roofline']['stage_ms_per_step']
print(sys.argv[2], 'resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'p50', round(d['latency_ms']['p50'],2), 'replay', round(s['replay'],2), 'kd', round(s['kd_order'],2), 'uf', round(s['union_find'],2), 'sort', round(s['component_sort'],2), 'parity', d['parity']['cluster_labels_equal_on_same_obstacle_cloud'] if d.get('parity') else None)" $1 "$2"; }
for cfg in "5 0 3" "5 1 3" "5 0 4" "5 0 2" "2 0 3"; do
  set -- $cfg
  LIDAR_B200_REPLAY_V=$1 LIDAR_B200_REPLAY5_LIVE=$2 LIDAR_B200_REPLAY5_CTAS_PER_SM=$3 timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sweep6_v$1_l$2_c$3.json 2> gpurun_out/sweep6.err
  summ gpurun_out/sweep6_v$1_l$2_c$3.json "v$1 live $2 ctas/sm $3:"
done 2>&1 | tee gpurun_out/sweep_replay_r2c6.txt
tail -3 gpurun_out/sweep6.err
LIDAR_B200_REPLAY_V=5 timeout 300 python tools/replay_stats.py > gpurun_out/replay_stats_v5.txt 2>&1; cat gpurun_out/replay_stats_v5.txt
for v in 5 2; do
  LIDAR_B200_REPLAY_V=$v timeout -k 10 300 python bench.py --workload merged1m --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_merged1m_v$v.json 2> gpurun_out/bench_merged1m.err
  summ gpurun_out/bench_merged1m_v$v.json "merged1m v$v:"
done
timeout 300 python tools/per_frame_stages.py > gpurun_out/per_frame_stages_v5.txt 2>&1; cat gpurun_out/per_frame_stages_v5.txt
