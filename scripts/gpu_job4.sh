mkdir -p gpurun_out
for g in 1.5 2; do
  for rows in 10000000 1250000; do
    FCS_TC_GROWTH=$g timeout 300 python bench.py --workload cfg3 --rows $rows --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('growth $g rows $rows: %.3f ms/step  %.0f q/s  K3 frac %.3f  fallbacks %d' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['tc_fallback_queries']))
"
  done
done
