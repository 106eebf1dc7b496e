#!/bin/bash
# round 2 evidence run (GPU box, 1 GPU): compute-sanitizer over the clustering kernels (incl. the window-synchronous
# replay), the default bench line with every extra, the reference arm, a 154-frame launch list and ncu --set full
# captures of the dominant kernels. Everything lands in gpurun_out/; the summaries are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
bash tools/gpu_sanitize_cluster.sh
timeout -k 10 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo "bench exit: $?"; tail -c 600 gpurun_out/bench_r2_final.json; tail -3 gpurun_out/bench_r2_final.err
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_ref.json 2> gpurun_out/bench_r2_final_ref.err
echo "ref exit: $?"; cut -c1-400 gpurun_out/bench_r2_final_ref.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_final.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_final.log 2>&1
echo "launches exit: $?"; wc -l gpurun_out/launches_r2_final.csv
for k in replay_gen_kernel replay_kernel cc_sample_kernel cc_link_kernel rs_scatter_kernel kd_level_kernel rs_frame_sort_kernel; do
  timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:"^(void )?(lb::)?${k}" -s 0 -c 1 \
     -f -o gpurun_out/r2final_${k} python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_${k}.log 2>&1
  echo "ncu $k exit: $?"
done
ls -la gpurun_out/r2final_*.ncu-rep
