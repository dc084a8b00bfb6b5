#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tc/group/fullsize"; timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_group_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -4
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra $BARGS 2>gpurun_out/err_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label: %.3f ms/step  %.0f q/s  e2e %.0f K3 frac %.3f  fallbacks %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['tc_fallback_queries']))
except Exception as e:
    print('$label: FAILED', e)"
  grep "fcs_tc\] phase" gpurun_out/err_$label.log | tail -${PH:-5} 
}
BARGS="--workload cfg3 --steps 10 --warmup 3"
PH=17 run cfg3 FCS_TC_PHASES=1
PH=0 run cfg3_nophase X=1
BARGS="--workload cfg3 --nq 512 --steps 20 --warmup 3"
PH=17 run nq512 FCS_TC_PHASES=1
PH=0 run nq512_nophase X=1
BARGS="--workload cfg3 --rows 1250000 --steps 20 --warmup 3"
PH=15 run 1.25M FCS_TC_PHASES=1
