#!/bin/bash
# round 2, GPU call 7: window-synchronous replay after the candidate-phase rework - parity, sweep, per-job phase stats,
# ncu full capture with source counters.
set -u
mkdir -p gpurun_out
STEPS="tests" PYTEST_ARGS="-x" PYTEST_TIMEOUT=900 bash tools/gpu_check.sh
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['roofline']['stage_ms_per_step']
print(sys.argv[2], 'resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'p50', round(d['latency_ms']['p50'],2), 'replay', round(s['replay'],2), 'kd', round(s['kd_order'],2), 'uf', round(s['union_find'],2), 'sort', round(s['component_sort'],2), 'parity', d['parity']['cluster_labels_equal_on_same_obstacle_cloud'] if d.get('parity') else None)" $1 "$2"; }
for cfg in "5 0 3" "5 0 2" "5 1 2"; do
  set -- $cfg
  LIDAR_B200_REPLAY_V=$1 LIDAR_B200_REPLAY5_LIVE=$2 LIDAR_B200_REPLAY5_CTAS_PER_SM=$3 timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sweep7_v$1_l$2_c$3.json 2> gpurun_out/sweep7.err
  summ gpurun_out/sweep7_v$1_l$2_c$3.json "v$1 live $2 ctas/sm $3:"
done 2>&1 | tee gpurun_out/sweep_replay_r2c7.txt
tail -3 gpurun_out/sweep7.err
LIDAR_B200_REPLAY_V=5 timeout 300 python tools/replay_stats.py > gpurun_out/replay_stats_v5b.txt 2>&1; cat gpurun_out/replay_stats_v5b.txt
LIDAR_B200_REPLAY_V=5 timeout -k 10 300 python bench.py --workload merged1m --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_merged1m_v5b.json 2> gpurun_out/bench_merged1m.err
summ gpurun_out/bench_merged1m_v5b.json "merged1m v5:"
STEPS="ncu" NCU_KERNELS="replay_gen_kernel" NCU_SKIP=0 NCU_BENCH_ARGS="--no-extras" bash tools/gpu_check.sh
