#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tc/group"; timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_group_gpu.py -m gpu -x -q 2>&1 | tail -3
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra $BARGS 2>gpurun_out/err_$label.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
    print('$label: %.3f ms/step  %.0f q/s  e2e %.0f K3 frac %.3f (step %.3f) fallbacks %d parity %s %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('step_frac',0), d['config']['tc_fallback_queries'], d['parity_checked'], d['parity']['errors'][:2]))
except Exception as e:
    print('$label: FAILED', e)"
  tail -3 gpurun_out/err_$label.log | cut -c1-300
}
BPOL="FCS_LIB_VARIANT=bpol FCS_TC_BPOLICY=1"
for rep in 1 2; do
BARGS="--workload cfg3 --steps 20 --warmup 3"
run base_cfg3_$rep X=1
run bpol_cfg3_$rep $BPOL
done
BARGS="--workload cfg3 --nq 512 --steps 20 --warmup 3"
run base_nq512 X=1
run bpol_nq512 $BPOL
BARGS="--workload cfg4b --steps 5 --warmup 3"
run base_cfg4b X=1
run bpol_cfg4b $BPOL
echo "== default bench line (all extras)"
( time timeout 900 python bench.py > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','parity_checked','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_burst'], d['cpu_baseline'])
for k,v in d['extra'].items():
    print(k, {kk: v.get(kk) for kk in ('value','ms_per_step','parity_checked','error')}, v.get('roofline',{}).get('frac') if isinstance(v.get('roofline'),dict) else None, {kk: v.get(kk) for kk in ('file_gbs','memmap_blocks_gbs')} if k=='loader' else '')
PY
