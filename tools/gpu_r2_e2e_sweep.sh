#!/bin/bash
# GPU box: e2e (lidar_b200_pipe_*) against the chunk size and pipeline depth, 1 GPU.  SWEEP="chunk:depth ..."
set -u
mkdir -p gpurun_out
: > gpurun_out/e2e_sweep.txt
for cd in ${SWEEP:-22:6 22:4 14:8 31:5 39:4 11:10 77:3}; do
  c=${cd%%:*}; d=${cd##*:}
  timeout -k 5 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --chunk $c --depth $d 2> /dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('chunk $c depth $d: resident %.0f e2e %.0f pageable %.0f p50 %.2f' % (d['value'], d['e2e']['value'], d['e2e']['pageable_host_buffers_value'], d['latency_ms']['p50']))
" >> gpurun_out/e2e_sweep.txt
done
cat gpurun_out/e2e_sweep.txt
