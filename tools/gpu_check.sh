#!/bin/bash
# Runs on the GPU box (gpurun): parity tests, smoke, bench, optional ncu captures.
#   STEPS: space-separated subset of "tests smoke bench ref launches ncu"   (default: tests smoke bench)
set -u
mkdir -p gpurun_out
STEPS=${STEPS:-"tests smoke bench"}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
for s in $STEPS; do
case $s in
tests)
  timeout -k 10 ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -q --timeout 900 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log ;;
smoke)
  timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
  echo "smoke exit: $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log ;;
bench)
  timeout -k 10 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit: $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
ref)
  timeout -k 10 900 python bench.py --impl reference ${BENCH_ARGS:-} > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "bench ref exit: $?"; tail -c 1500 gpurun_out/bench_ref.json ;;
launches)
  timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline ${NCU_BENCH_ARGS:---frames 16} > gpurun_out/ncu_bench.log 2>&1
  echo "launches exit: $?"; wc -l gpurun_out/launches.csv ;;
ncu)
  for k in ${NCU_KERNELS:-replay_cta_kernel}; do
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"${k}" -s ${NCU_SKIP:-1} -c ${NCU_COUNT:-1} \
     -f -o gpurun_out/prof_${k} python bench.py --steps 1 --warmup 1 --no-cpu-baseline ${NCU_BENCH_ARGS:---frames 16} > gpurun_out/ncu_full_${k}.log 2>&1
  echo "ncu $k exit: $?"
  done
  ls -la gpurun_out/*.ncu-rep ;;
esac
done
