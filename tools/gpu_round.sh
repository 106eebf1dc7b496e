#!/bin/bash
# GPU box: parity tests, smoke, the default bench line, the three synthetic-shape lines and a 154-frame launch list.
set -u
mkdir -p gpurun_out
STEPS="tests smoke bench ref" bash tools/gpu_check.sh
for w in synth128 merged1m synth64; do
  timeout -k 10 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w exit: $?"; tail -c 1200 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_154.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_154.log 2>&1
echo "launches154 exit: $?"; wc -l gpurun_out/launches_154.csv
