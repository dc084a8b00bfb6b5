#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tc"; timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_group_gpu.py -m gpu -x -q 2>&1 | tail -4
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra $BARGS 2>gpurun_out/err_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label: %.3f ms/step  %.0f q/s  e2e %.0f  K3 frac %.3f  fallbacks %d launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['tc_fallback_queries'], d['gpu_launches']))
except Exception as e:
    print('$label: FAILED', e)"
  grep "fcs_tc\] phase" gpurun_out/err_$label.log | tail -${PH:-14}
}
BARGS="--workload cfg3 --steps 10 --warmup 3"
PH=14 run base FCS_TC_PHASES=1
PH=0 run base2 X=1
PH=0 run r0_8 FCS_TC_R0_TILES=8
PH=0 run cs512 FCS_TC_C_SAMPLE=512
PH=0 run cs1024m8 FCS_TC_C_SAMPLE=1024 FCS_TC_M_MIN=8
PH=0 run cf512 FCS_TC_CF_MIN=512 FCS_TC_CF_MULT=3
BARGS="--workload cfg3 --nq 512 --steps 20 --warmup 3"
PH=12 run nq512 FCS_TC_PHASES=1
PH=0 run nq512_r0_8 FCS_TC_R0_TILES=8
PH=0 run nq512_cs512 FCS_TC_C_SAMPLE=512 FCS_TC_R0_TILES=8
PH=0 run nq512_cs1024m8 FCS_TC_C_SAMPLE=1024 FCS_TC_M_MIN=8 FCS_TC_R0_TILES=8
BARGS="--workload cfg3 --rows 1250000 --steps 20 --warmup 3"
PH=12 run rows1.25M FCS_TC_PHASES=1
PH=0 run rows1.25M_r0_8 FCS_TC_R0_TILES=8
BARGS="--workload cfg4b --steps 5 --warmup 3"
PH=0 run cfg4b X=1
