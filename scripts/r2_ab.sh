#!/bin/bash
# A/B of two library builds on one box: default vs the variant named in $VAR (env assignments)
mkdir -p gpurun_out
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra $BARGS 2>gpurun_out/err_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label: %.3f ms/step  %.0f q/s  e2e %.0f K3 frac %.3f fallbacks %d parity %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['run']['tc_fallback_queries'], d['parity_checked']))
except Exception as e:
    print('$label: FAILED', e)"
}
for rep in 1 2; do
BARGS="--workload cfg3 --steps 40 --warmup 3"
run base_cfg3_$rep X=1
run var_cfg3_$rep $VAR
BARGS="--workload cfg4b --steps 20 --warmup 3"
run base_cfg4b_$rep X=1
run var_cfg4b_$rep $VAR
done
