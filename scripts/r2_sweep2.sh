#!/bin/bash
mkdir -p gpurun_out
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra $BARGS 2>gpurun_out/err_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label: %.3f ms/step  %.0f q/s  e2e %.0f K3 frac %.3f fallbacks %d rounds %d parity %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['tc_fallback_queries'], d['config']['tc_rounds'], d['parity_checked']))
except Exception as e:
    print('$label: FAILED', e)"
}
for rep in 1 2; do
BARGS="--workload cfg3 --steps 40 --warmup 3"
run base_$rep X=1
run cf512_$rep FCS_TC_CF_MIN=512 FCS_TC_CF_MULT=3.2
run cf640_$rep FCS_TC_CF_MIN=640 FCS_TC_CF_MULT=4.0
run cs256_$rep FCS_TC_C_SAMPLE=256
run cs64_$rep FCS_TC_C_SAMPLE=64 FCS_TC_M_MIN=8
done
BARGS="--workload cfg3 --nq 512 --steps 60 --warmup 3"
run nq512_base X=1
run nq512_cf512 FCS_TC_CF_MIN=512 FCS_TC_CF_MULT=3.2
run nq512_cs256 FCS_TC_C_SAMPLE=256
BARGS="--workload cfg3 --rows 1250000 --steps 60 --warmup 3"
run 1.25M_base X=1
run 1.25M_cf256 FCS_TC_CF_MIN=256 FCS_TC_CF_MULT=1.6
