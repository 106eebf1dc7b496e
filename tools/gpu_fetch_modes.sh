#!/bin/bash
# e2e throughput of the frame pipeline per LIDAR_B200_FETCH_MODE and pipeline depth (GPU box)
set -u
mkdir -p gpurun_out
CFGS=${CFGS:-0:6 3:6 3:8 3:10 3:12}
for cfg in $CFGS; do
  m=${cfg%%:*}; d=${cfg##*:}
  LIDAR_B200_FETCH_MODE=$m timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --depth $d ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('mode $m depth $d: resident', round(d['value']), 'e2e', round(e['value']), 'pageable', round(e['pageable_host_buffers_value']), 'd2h MB', round(e['d2h_bytes_per_step']/1e6), 'equal', e['results_equal_resident_run'])" | tee -a gpurun_out/fetch_modes.log
done
