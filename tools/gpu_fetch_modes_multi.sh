#!/bin/bash
# e2e throughput at N GPUs per LIDAR_B200_FETCH_MODE:  bash tools/gpu_fetch_modes_multi.sh <N> "<modes>"
set -u
mkdir -p gpurun_out
n=${1:-4}
for m in ${2:-0 3}; do
  LIDAR_B200_FETCH_MODE=$m timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); e=d['e2e']
print('gpus $n mode $m: resident', round(d['value']), 'e2e', round(e['value']), 'pageable', round(e['pageable_host_buffers_value']), 'd2h MB', round(e['d2h_bytes_per_step']/1e6), 'equal', e['results_equal_resident_run'])" | tee -a gpurun_out/fetch_modes_multi.log
done
