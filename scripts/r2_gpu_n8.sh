#!/bin/bash
# 8 GPUs: NCCL parity tests, scaling lines N=1/2/4/8 (default bench incl. cfg4/cfg4b/cfg5 extras), row-sharded cfg3, local engine
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== pytest multigpu + group"; timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_group_gpu.py -m gpu -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { python -c "
import json,sys
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1])
    print('$1: n=%d %.3f ms/step %.0f q/s e2e %.0f frac %.3f parity %s fb %s | %s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_checked'], d['run']['tc_fallback_queries'], d['config']['parallelism']))
    for k,v in d.get('extra',{}).items(): print('   ', k, v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('value'), v.get('parity_checked'), (v.get('roofline') or {}).get('frac'), v.get('error'))
except Exception as e:
    print('$1 FAILED', e)
" ; }
for n in 8 4 2; do
echo "== bench N=$n auto layout"; ( time timeout 1200 $TR --nproc-per-node $n --master-port 2966$n bench.py --gpus $n > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err ) 2>&1 | grep real; show gpurun_out/bench_n$n.json
done
echo "== bench N=1"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; show gpurun_out/bench_n1.json
for n in 8 4 2; do
echo "== bench N=$n row-sharded cfg3"; timeout 900 $TR --nproc-per-node $n --master-port 2967$n bench.py --gpus $n --query-groups 1 --no-extra > gpurun_out/bench_n${n}_q1.json 2> gpurun_out/bench_n${n}_q1.err; show gpurun_out/bench_n${n}_q1.json
done
echo "== cfg2 N=8 (strong, 500k rows)"; timeout 600 $TR --nproc-per-node 8 --master-port 29681 bench.py --gpus 8 --workload cfg2 --no-extra > gpurun_out/bench_n8_cfg2.json 2> gpurun_out/bench_n8_cfg2.err; show gpurun_out/bench_n8_cfg2.json
echo "== local engine, 8 GPUs"; timeout 1200 python bench.py --workload local > gpurun_out/bench_local_n8.json 2> gpurun_out/bench_local_n8.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_local_n8.json').read().strip().splitlines()[-1])
for c in d['cases']: print(c)"; tail -3 gpurun_out/bench_local_n8.err | cut -c1-300
echo "== reference arm N=8"; timeout 600 $TR --nproc-per-node 8 --master-port 29682 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
