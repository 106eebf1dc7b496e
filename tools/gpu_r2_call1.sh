#!/bin/bash
# round 2, GPU call 1: stricter parity tests, sanitizers over the union-find / replay kernels, baseline bench (all
# workloads + config 5), launch list and ncu --set full of the union-find / sort / replay kernels at 154 frames,
# copy-only probe of the e2e byte pattern.
set -u
mkdir -p gpurun_out
STEPS="tests smoke" bash tools/gpu_check.sh
SAN_TIMEOUT=420 bash tools/gpu_sanitize_cluster.sh
timeout -k 10 900 python bench.py > gpurun_out/bench_r2_base.json 2> gpurun_out/bench_r2_base.err
echo "bench exit: $?"; tail -c 2500 gpurun_out/bench_r2_base.json; tail -5 gpurun_out/bench_r2_base.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err
echo "ref exit: $?"; cut -c1-300 gpurun_out/bench_r2_ref.json
for bpp in 16 9.7; do timeout 120 python tools/copy_probe.py --d2h-bytes-per-point $bpp; done > gpurun_out/copy_probe_1gpu.json 2>&1
cat gpurun_out/copy_probe_1gpu.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_base.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches.log 2>&1
echo "launches exit: $?"; wc -l gpurun_out/launches_r2_base.csv
for k in cc_sample_kernel cc_link_kernel rs_scatter replay_cta_kernel replay_kernel; do
  timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:"${k}" -s 1 -c 1 \
     -f -o gpurun_out/r2base_${k} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_${k}.log 2>&1
  echo "ncu $k exit: $?"
done
ls -la gpurun_out/*.ncu-rep
