#!/bin/bash
# e2e pipeline sweep on the GPU box: chunk size x depth
for cd in "77 2" "39 3" "39 4" "22 3" "22 4" "22 6" "11 6" "11 8"; do
  set -- $cd
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chunk $1 --depth $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('chunk $1 depth $2: e2e', round(e['value']), 'pageable', round(e['pageable_host_buffers_value']), 'resident', round(d['value']), 'equal', e['results_equal_resident_run'], 'p50', d['latency_ms']['p50'])"
done
