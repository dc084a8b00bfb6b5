#!/bin/bash
# Round 2, final 1-GPU run: full GPU test suite, smoke, ncu evidence, bench lines for profiles/
mkdir -p gpurun_out
echo "== pytest (all gpu tests)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
bash scripts/r2_evidence.sh 2>&1 | tail -4
echo "== phases"; FCS_TC_PHASES=1 timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>&1 >/dev/null | grep "fcs_tc\] phase" | tail -17 > gpurun_out/r02_phases_cfg3.txt; cat gpurun_out/r02_phases_cfg3.txt
echo "== default bench line"; ( time timeout 1200 python bench.py > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err ) 2>&1 | grep real
for w in cfg2 cfg4 cfg4b cfg5; do
  timeout 900 python bench.py --workload $w --no-extra > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_cfg3.json 2>/dev/null
python - <<'PY'
import json
for w in ('default_n1','cfg2','cfg4','cfg4b','cfg5'):
    try:
        d=json.loads(open(f'gpurun_out/r02_bench_{w}.json').read().strip().splitlines()[-1])
        print(w, 'value %.1f' % d['value'], 'ms %.4f' % d['ms_per_step'], 'e2e %.1f' % d['e2e']['value'], 'frac %.3f' % d['roofline']['frac'], 'parity', d['parity_checked'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))
        for k,v in d.get('extra',{}).items():
            print('   ', k, v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('value'), v.get('parity_checked'), (v.get('roofline') or {}).get('frac'), v.get('error'), {kk: v.get(kk) for kk in ('file_gbs','memmap_blocks_gbs')} if k=='loader' else '')
    except Exception as e:
        print(w, 'FAILED', e)
PY
