#!/bin/bash
mkdir -p gpurun_out
echo "== pytest (all gpu tests)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== default bench line"
( time timeout 1200 python bench.py > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_default_n1.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','parity_checked','gpu_launches')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['frac_burst'], 'cpu', d['cpu_baseline']['value'])
for k,v in d['extra'].items():
    print(k, {kk: v.get(kk) for kk in ('value','ms_per_step','parity_checked','error')}, v.get('roofline',{}).get('frac') if isinstance(v.get('roofline'),dict) else None, (v.get('e2e') or {}).get('value'), {kk: v.get(kk) for kk in ('file_gbs','memmap_blocks_gbs')} if k=='loader' else '')
PY
