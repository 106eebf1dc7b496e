STEPS="launches" bash tools/gpu_check.sh
for c in 4 8 16; do
  echo "== replay ctas/sm $c"
  LIDAR_B200_REPLAY_CTAS_PER_SM=$c python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
done
