#!/bin/bash
# generic sweep on the GPU box: SWEEP_VAR=<env var> SWEEP_VALUES="a b c" [BENCH_ARGS=...] bash tools/sweep_env.sh
mkdir -p gpurun_out
for v in $SWEEP_VALUES; do
  env $SWEEP_VAR=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; st=d['roofline']['stage_ms_per_step']
print('$SWEEP_VAR=$v: resident', round(d['value']), 'e2e', round(e['value']), 'p50', round(d['latency_ms']['p50'],2), 'parity', d.get('parity') and {k: v for k, v in d['parity'].items() if k != 'what'}, 'stages', {k: round(x,2) for k,x in st.items()})" | tee -a gpurun_out/sweep.log
done
