#!/bin/bash
# round 2, GPU call 3: CTA replay with candidates dealt over the whole CTA - parity, A/B, CTA-threshold sweep, round
# statistics, ncu; result fetch through emit_results_kernel (mode 4) vs copies (mode 0); the real drop-in call; per-frame stages.
set -u
mkdir -p gpurun_out
STEPS="tests smoke" bash tools/gpu_check.sh
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['roofline']['stage_ms_per_step']
print(sys.argv[2], 'resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'p50', round(d['latency_ms']['p50'],2), 'replay', round(s['replay'],2), 'uf', round(s['union_find'],2), 'd2h', d['e2e']['d2h_bytes_per_step'], 'parity', d['parity']['cluster_labels_equal_on_same_obstacle_cloud'] if d.get('parity') else None)" $1 "$2"; }
for cfg in "2 256" "1 256" "2 512" "2 1024" "2 2048" "2 4096" "2 1000000000" "1 1024" "1 4096"; do
  set -- $cfg
  LIDAR_B200_REPLAY_V=$1 LIDAR_B200_CTA_MIN_MEMBERS=$2 timeout -k 10 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sweep_v$1_m$2.json 2> gpurun_out/sweep.err
  summ gpurun_out/sweep_v$1_m$2.json "v$1 cta_min $2:"
done 2>&1 | tee gpurun_out/sweep_replay_r2.txt
LIDAR_B200_FETCH_MODE=4 timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_fetch4.json 2> gpurun_out/bench_fetch4.err
summ gpurun_out/bench_fetch4.json "fetch mode 4:"; tail -2 gpurun_out/bench_fetch4.err
LIDAR_B200_REPLAY_STATS=1 timeout 300 python tools/replay_stats.py > gpurun_out/replay_stats_v2b.txt 2>&1; cat gpurun_out/replay_stats_v2b.txt
timeout 300 python - > gpurun_out/dropin_timing.json 2> gpurun_out/dropin_timing.err <<'PY'
import json, sys
sys.path.insert(0, '.')
import bench
frames, _ = bench.load_workload()
print(json.dumps(bench.dropin_timing(frames)))
PY
cat gpurun_out/dropin_timing.json; tail -3 gpurun_out/dropin_timing.err
timeout 300 python tools/per_frame_stages.py > gpurun_out/per_frame_stages.txt 2>&1; cat gpurun_out/per_frame_stages.txt
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:replay_cta2_kernel -s 0 -c 1 \
   -f -o gpurun_out/r2_replay_cta2b python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_replay_cta2b.log 2>&1
echo "ncu exit: $?"
