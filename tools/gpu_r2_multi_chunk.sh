#!/bin/bash
# GPU box with N GPUs: the e2e pipeline with the old (22 x 6) and new (77 x 4) chunking under torchrun.  N=${N:-8}
set -u
mkdir -p gpurun_out
N=${N:-8}
: > gpurun_out/multi_chunk_${N}gpu.txt
for cd in 22:6 77:4; do
  c=${cd%%:*}; d=${cd##*:}
  timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
     bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-extras --chunk $c --depth $d 2> gpurun_out/multi_chunk.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$N gpus chunk $c depth $d: resident %.0f e2e %.0f fetch mode %s p50 %.2f' % (d['value'], d['e2e']['value'], d['e2e'].get('fetch_mode'), d['latency_ms']['p50']))
" >> gpurun_out/multi_chunk_${N}gpu.txt
done
cat gpurun_out/multi_chunk_${N}gpu.txt; tail -2 gpurun_out/multi_chunk.err
