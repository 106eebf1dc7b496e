#!/bin/bash
# round 2, GPU call 2: second-generation CTA replay (state in shared memory) - parity, A/B against the first generation,
# per-job round statistics, ncu --set full of the new kernel.
set -u
mkdir -p gpurun_out
STEPS="tests smoke" bash tools/gpu_check.sh
for v in 2 1; do
  LIDAR_B200_REPLAY_V=$v timeout -k 10 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_replay_v$v.json 2> gpurun_out/bench_replay_v$v.err
  echo "bench v$v exit: $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_replay_v$v.json')); print('v$v', d['value'], d['e2e']['value'], d['latency_ms']['p50'], d['roofline']['stage_ms_per_step'], d['parity'])"
  LIDAR_B200_REPLAY_V=$v timeout -k 10 600 python bench.py --workload merged1m --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_merged1m_v$v.json 2> gpurun_out/bench_merged1m_v$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_merged1m_v$v.json')); print('merged1m v$v', d['value'], d['e2e']['value'], d['roofline']['stage_ms_per_step'])"
done
LIDAR_B200_REPLAY_STATS=1 timeout 300 python tools/replay_stats.py > gpurun_out/replay_stats_v2.txt 2>&1; cat gpurun_out/replay_stats_v2.txt
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:replay_cta2_kernel -s 1 -c 1 \
   -f -o gpurun_out/r2_replay_cta2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_replay_cta2.log 2>&1
echo "ncu exit: $?"
