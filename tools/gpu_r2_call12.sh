#!/bin/bash
# round 2, GPU call 12: wide CTAs (512 / 1024 threads) for the long jobs of the window-synchronous replay.
set -u
mkdir -p gpurun_out
STEPS="tests" PYTEST_ARGS="-x" PYTEST_TIMEOUT=900 bash tools/gpu_check.sh
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['roofline']['stage_ms_per_step']
print(sys.argv[2], 'resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'p50', round(d['latency_ms']['p50'],2), {k: round(v,2) for k,v in s.items()}, 'parity', d['parity']['cluster_labels_equal_on_same_obstacle_cloud'] if d.get('parity') else None)" $1 "$2"; }
for cfg in "512 3" "1024 3" "256 3" "512 2" "1024 2"; do
  set -- $cfg
  LIDAR_B200_REPLAY5_BIG_THREADS=$1 LIDAR_B200_REPLAY5_CTAS_PER_SM=$2 timeout -k 10 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sweep12_t$1_c$2.json 2> gpurun_out/sweep12.err
  summ gpurun_out/sweep12_t$1_c$2.json "big threads $1 small ctas/sm $2:"
done 2>&1 | tee gpurun_out/sweep_r2c12.txt
tail -3 gpurun_out/sweep12.err
for t in 512 1024; do
LIDAR_B200_REPLAY5_BIG_THREADS=$t python tools/replay_stats_single.py 143 0 > gpurun_out/replay_single_t$t.txt 2>&1; head -5 gpurun_out/replay_single_t$t.txt
done
for t in 512 1024 256; do
  LIDAR_B200_REPLAY5_BIG_THREADS=$t timeout -k 10 300 python bench.py --workload merged1m --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_merged1m_t$t.json 2> gpurun_out/bench_merged1m.err
  summ gpurun_out/bench_merged1m_t$t.json "merged1m big threads $t:"
done
