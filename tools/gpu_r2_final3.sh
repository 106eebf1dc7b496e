#!/bin/bash
# round 2, final evidence part 3 (1 GPU): pipeline tests + default bench line with the 77 x 4 pipeline
set -u
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 600 -k "pipeline or pipe or e2e or multi" > gpurun_out/pytest_pipe.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_pipe.log; tail -3 gpurun_out/pytest_pipe.log
timeout -k 10 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err
echo "bench exit: $?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_r2_final.json'))
print(d['value'], d['e2e']['value'], d['e2e']['pageable_host_buffers_value'], d['latency_ms']['p50'], d['e2e'].get('what','')[:80])
for k, v in (d.get('workloads') or {}).items():
    print(k, v.get('value'), v.get('e2e'))
print(d['config5']['value'], d['config5']['e2e'])
PY
tail -3 gpurun_out/bench_r2_final.err
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_final_ref.json 2> gpurun_out/bench_r2_final_ref.err
echo "ref exit: $?"; cut -c1-200 gpurun_out/bench_r2_final_ref.json
