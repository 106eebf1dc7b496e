#!/bin/bash
# replay tuning sweep on the GPU box
for cfg in "256 2" "512 2" "1024 2" "2048 2" "4096 2" "1000000000 2" "256 3" "1024 3"; do
  set -- $cfg
  LIDAR_B200_CTA_MIN_MEMBERS=$1 LIDAR_B200_REPLAY_BIG_CTAS_PER_SM=$2 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('cta_min $1 ctas/sm $2: resident', round(d['value']), 'e2e', round(e['value']), 'p50', round(d['latency_ms']['p50'],2), 'replay ms', round(d['roofline']['stage_ms_per_step']['replay'],2))"
done
