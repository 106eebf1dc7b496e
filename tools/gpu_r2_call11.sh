#!/bin/bash
# round 2, GPU call 11: window latency work on the window-synchronous replay - parity, single-frame phase stats, bench.
set -u
mkdir -p gpurun_out
STEPS="tests" PYTEST_ARGS="-x" PYTEST_TIMEOUT=900 bash tools/gpu_check.sh
python tools/replay_stats_single.py 143 0 > gpurun_out/replay_single.txt 2>&1; cat gpurun_out/replay_single.txt
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['roofline']['stage_ms_per_step']
print(sys.argv[2], 'resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'fetch', d['e2e'].get('fetch_mode'), 'p50', round(d['latency_ms']['p50'],2), {k: round(v,2) for k,v in s.items()}, 'parity', d['parity']['cluster_labels_equal_on_same_obstacle_cloud'] if d.get('parity') else None)" $1 "$2"; }
for c in 3 2; do
LIDAR_B200_REPLAY5_CTAS_PER_SM=$c timeout -k 10 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sweep11_c$c.json 2> gpurun_out/sweep11.err
summ gpurun_out/sweep11_c$c.json "ctas/sm $c:"
done
tail -3 gpurun_out/sweep11.err
LIDAR_B200_REPLAY_V=5 timeout 300 python tools/replay_stats.py > gpurun_out/replay_stats_v5c.txt 2>&1; head -8 gpurun_out/replay_stats_v5c.txt
