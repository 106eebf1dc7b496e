#!/bin/bash
# round 2, GPU call 9: frame-resident radix sort (one CTA per frame through all passes) - parity, A/B against the
# per-tile kernels; the pipe's own choice of the result fetch mode at 1 GPU.
set -u
mkdir -p gpurun_out
STEPS="tests smoke" PYTEST_ARGS="-x" PYTEST_TIMEOUT=900 bash tools/gpu_check.sh
summ() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); s=d['roofline']['stage_ms_per_step']
print(sys.argv[2], 'resident', round(d['value']), 'e2e', round(d['e2e']['value']), 'fetch', d['e2e'].get('fetch_mode'), 'p50', round(d['latency_ms']['p50'],2), {k: round(v,2) for k,v in s.items()}, 'parity', d['parity']['cluster_labels_equal_on_same_obstacle_cloud'] if d.get('parity') else None)" $1 "$2"; }
for fs in 1 0; do
  LIDAR_B200_FRAME_SORT=$fs timeout -k 10 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/sweep9_fs$fs.json 2> gpurun_out/sweep9.err
  summ gpurun_out/sweep9_fs$fs.json "frame sort $fs:"
done 2>&1 | tee gpurun_out/sweep_r2c9.txt
tail -3 gpurun_out/sweep9.err
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_c9.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches9.log 2>&1
echo "launches exit: $?"; grep -c rs_frame_sort gpurun_out/launches_r2_c9.csv
