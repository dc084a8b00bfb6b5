#!/bin/bash
mkdir -p gpurun_out
EPI="FCS_LIB_VARIANT=vote FCS_TC_VOTE=1"
echo "== tests (default build: hit mask)"; timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_group_gpu.py -m gpu -x -q 2>&1 | tail -3
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra $BARGS 2>gpurun_out/err_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label: %.3f ms/step  %.0f q/s  e2e %.0f K3 frac %.3f fallbacks %d parity %s clk %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['tc_fallback_queries'], d['parity_checked'], d['clocks']['sm_mhz']))
except Exception as e:
    print('$label: FAILED', e)"
  grep "fcs_tc\] phase gemm" gpurun_out/err_$label.log | tail -${PH:-0} | tr '\n' ' '; echo
}
for rep in 1 2; do
BARGS="--workload cfg3 --steps 40 --warmup 3"
run base_cfg3_$rep X=1
run epi_cfg3_$rep $EPI
BARGS="--workload cfg4b --steps 20 --warmup 3"
run base_cfg4b_$rep X=1
run epi_cfg4b_$rep $EPI
done
BARGS="--workload cfg3 --nq 512 --steps 60 --warmup 3"
run base_nq512 X=1
run epi_nq512 $EPI
BARGS="--workload cfg3 --rows 1250000 --steps 60 --warmup 3"
run base_1.25M X=1
run epi_1.25M $EPI
BARGS="--workload cfg3 --steps 5 --warmup 3"
PH=5 run epi_phases $EPI FCS_TC_PHASES=1
