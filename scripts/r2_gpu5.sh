#!/bin/bash
# A/B on one box: slow path from registers (default build) vs 8-column TMEM re-read (variant ld8), alternating
mkdir -p gpurun_out
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra $BARGS 2>gpurun_out/err_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label: %.3f ms/step  %.0f q/s  K3 frac %.3f  fallbacks %d' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['tc_fallback_queries']))
except Exception as e:
    print('$label: FAILED', e)"
  grep "fcs_tc\] phase gemm" gpurun_out/err_$label.log | tail -${PH:-5} | tr '\n' ' '; echo
}
LD8="FCS_LIB_VARIANT=ld8 FCS_TC_SLOWPATH=1"
for rep in 1 2; do
BARGS="--workload cfg3 --steps 10 --warmup 3"
PH=5 run reg_cfg3_$rep FCS_TC_PHASES=1
PH=5 run ld8_cfg3_$rep $LD8 FCS_TC_PHASES=1
BARGS="--workload cfg4b --steps 5 --warmup 3"
PH=6 run reg_cfg4b_$rep FCS_TC_PHASES=1
PH=6 run ld8_cfg4b_$rep $LD8 FCS_TC_PHASES=1
BARGS="--workload cfg3 --rows 1250000 --steps 20 --warmup 3"
PH=4 run reg_1.25M_$rep FCS_TC_PHASES=1
PH=4 run ld8_1.25M_$rep $LD8 FCS_TC_PHASES=1
BARGS="--workload cfg3 --nq 512 --steps 20 --warmup 3"
PH=5 run reg_nq512_$rep FCS_TC_PHASES=1
PH=5 run ld8_nq512_$rep $LD8 FCS_TC_PHASES=1
done
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv
