#!/bin/bash
mkdir -p gpurun_out
echo "== default bench line (all extras)"
( time timeout 1200 python bench.py > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err ) 2>&1 | tail -3
tail -5 gpurun_out/bench_default_n1.err | cut -c1-400
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','parity_checked','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_burst'], d['cpu_baseline']['value'])
print(d['parity'])
for k,v in d['extra'].items():
    print(k, {kk: v.get(kk) for kk in ('value','ms_per_step','parity_checked','error')}, v.get('roofline',{}).get('frac') if isinstance(v.get('roofline'),dict) else None, (v.get('e2e') or {}).get('value'), {kk: v.get(kk) for kk in ('file_gbs','memmap_blocks_gbs')} if k=='loader' else '', v.get('parity',{}).get('errors') if isinstance(v.get('parity'),dict) else '')
PY
echo "== reference arm"; ( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) 2>&1 | tail -5 | cut -c1-600
echo "== cfg5 slice"; timeout 600 python bench.py --workload cfg5 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5 slice: %.1f ms/step %.0f q/s e2e %.0f frac %.3f parity %s fb %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_checked'], d['config']['tc_fallback_queries']))"
