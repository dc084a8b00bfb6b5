#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tc/group/gemv"; timeout 1200 python -m pytest tests/test_tc_gpu.py tests/test_group_gpu.py tests/test_gemv_gpu.py -m gpu -x -q 2>&1 | tail -12
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra $BARGS 2>gpurun_out/err_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label: %.3f ms/step  %.0f q/s  e2e %.0f K3 frac %.3f (step %.3f) fallbacks %d parity %s %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('step_frac',0), d['config']['tc_fallback_queries'], d['parity_checked'], d['parity']['errors'][:2]))
except Exception as e:
    print('$label: FAILED', e)"
  tail -2 gpurun_out/err_$label.log | cut -c1-300
}
BARGS="--workload cfg3 --steps 20 --warmup 3"
run cfg3 X=1
