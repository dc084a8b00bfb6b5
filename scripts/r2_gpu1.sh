#!/bin/bash
# Round 2, GPU call 1 (1 GPU): parity of the re-designed tensor-core path, then the plan-parameter sweep on cfg3.
mkdir -p gpurun_out
echo "== pytest tc + gemv + dropin"; timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_gemv_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -15
echo "== pytest fullsize"; timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -8
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
run() {  # label, env..., -- bench args
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra $BARGS 2>gpurun_out/err_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label: %.3f ms/step  %.0f q/s  e2e %.0f  K3 frac %.3f  fallbacks %d launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['tc_fallback_queries'], d['gpu_launches']))
except Exception as e:
    print('$label: FAILED', e)"
  grep "fcs_tc\] round" gpurun_out/err_$label.log | tail -8
}
BARGS="--workload cfg3 --steps 10 --warmup 3"
run base FCS_TC_VERBOSE=1
run cf4 FCS_TC_CF_MULT=4 FCS_TC_CF_MIN=640
run cfmin384 FCS_TC_CF_MIN=384 FCS_TC_M_FINAL=24
run m16 FCS_TC_M_FINAL=16
run m32 FCS_TC_M_FINAL=32
run m48 FCS_TC_M_FINAL=48
run cs512 FCS_TC_C_SAMPLE=512
run cs2048 FCS_TC_C_SAMPLE=2048 FCS_TC_M_MIN=16
BARGS="--workload cfg3 --nq 512 --steps 20 --warmup 3"
run nq512 FCS_TC_VERBOSE=1
run nq512_m48 FCS_TC_M_FINAL=48
BARGS="--workload cfg3 --rows 1250000 --steps 20 --warmup 3"
run rows1.25M FCS_TC_VERBOSE=1
run rows1.25M_m48 FCS_TC_M_FINAL=48
BARGS="--workload cfg4b --steps 5 --warmup 3"
run cfg4b FCS_TC_VERBOSE=1
BARGS="--workload cfg2 --steps 2000 --warmup 10"
run cfg2 X=1
