#!/bin/bash
# GPU box: wide (512-thread) CTAs for the longest replay jobs in BATCH mode: which lists, how many CTAs per SM
set -u
mkdir -p gpurun_out
: > gpurun_out/replay_wide.txt
run() {
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); st = d['roofline']['stage_ms_per_step']; print('$*:', 'resident %.0f e2e %.0f p50 %.2f replay %.2f uf %.2f kd %.2f' % (d['value'], d['e2e']['value'], d['latency_ms']['p50'], st['replay'], st['union_find'], st['kd_order']))
" >> gpurun_out/replay_wide.txt
}
run LIDAR_B200_REPLAY5_BIG_THREADS=256
run LIDAR_B200_REPLAY5_BIG_THREADS=512 LIDAR_B200_REPLAY5_WIDE_LISTS=2 LIDAR_B200_REPLAY5_WIDE_CTAS_PER_SM=2
run LIDAR_B200_REPLAY5_BIG_THREADS=512 LIDAR_B200_REPLAY5_WIDE_LISTS=1 LIDAR_B200_REPLAY5_WIDE_CTAS_PER_SM=2
run LIDAR_B200_REPLAY5_BIG_THREADS=512 LIDAR_B200_REPLAY5_WIDE_LISTS=1 LIDAR_B200_REPLAY5_WIDE_CTAS_PER_SM=1
cat gpurun_out/replay_wide.txt
