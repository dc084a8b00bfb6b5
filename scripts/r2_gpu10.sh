#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tc/group"; timeout 1200 python -m pytest tests/test_tc_gpu.py tests/test_group_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -5
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra $BARGS 2>gpurun_out/err_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label: %.3f ms/step  %.0f q/s  e2e %.0f K3 frac %.3f (step %.3f) fallbacks %d rounds %d parity %s %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('step_frac',0), d['config']['tc_fallback_queries'], d['config']['tc_rounds'], d['parity_checked'], d['parity']['errors'][:2]))
except Exception as e:
    print('$label: FAILED', e)"
  tail -2 gpurun_out/err_$label.log | cut -c1-300
}
for rep in 1 2; do
BARGS="--workload cfg3 --steps 20 --warmup 3"
run pdl_cfg3_$rep X=1
run nopdl_cfg3_$rep FCS_TC_PDL=0
BARGS="--workload cfg3 --nq 512 --steps 40 --warmup 3"
run pdl_nq512_$rep X=1
run nopdl_nq512_$rep FCS_TC_PDL=0
run nopdl_nq512_cs128_$rep FCS_TC_PDL=0 FCS_TC_C_SAMPLE=128
BARGS="--workload cfg3 --rows 1250000 --steps 40 --warmup 3"
run pdl_1.25M_$rep X=1
run nopdl_1.25M_$rep FCS_TC_PDL=0
done
BARGS="--workload cfg4b --steps 5 --warmup 3"
run pdl_cfg4b X=1
