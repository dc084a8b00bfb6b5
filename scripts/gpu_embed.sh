mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_embed_gpu.py -x -q -s -k "tc" 2>&1 | tail -8
timeout 200 python bench.py --workload embed --no-cpu-baseline 2>/dev/null > gpurun_out/bench_embed_v4.json; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_embed_v4.json').read().strip().splitlines()[-1])
print('embed: %.1f structures/s  %.1f ms/step  e2e %.1f  edge kernel %.1f TFLOP/s = %.3f of peak (%s) at %s MHz' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['bound'], d['clocks']['sm_mhz']))"
