#!/bin/bash
# 2 GPUs: NCCL parity tests, group over 2 devices, bench at N=2 (both layouts), distributed_embed check, local engine
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest multigpu + group"; timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_group_gpu.py -m gpu -x -q 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== distributed_embed over NCCL"; timeout 600 $TR --nproc-per-node 2 --master-port 29655 scripts/dist_embed_check.py 2>&1 | tail -2
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1])
print('$1: n=%d %.3f ms/step %.0f q/s e2e %.0f frac %.3f parity %s | %s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_checked'], d['config']['parallelism']))
for k,v in d.get('extra',{}).items(): print('   ', k, v.get('value'), v.get('ms_per_step'), (v.get('e2e') or {}).get('value'), v.get('parity_checked'), v.get('error'))
" ; }
echo "== bench N=2 auto layout"; timeout 900 $TR --nproc-per-node 2 --master-port 29656 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; show gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err | cut -c1-300
echo "== bench N=2 row-sharded"; timeout 900 $TR --nproc-per-node 2 --master-port 29657 bench.py --gpus 2 --query-groups 1 --no-extra > gpurun_out/bench_n2_q1.json 2> gpurun_out/bench_n2_q1.err; show gpurun_out/bench_n2_q1.json
echo "== local engine"; timeout 900 python bench.py --workload local > gpurun_out/bench_local_n2.json 2> gpurun_out/bench_local_n2.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_local_n2.json').read().strip().splitlines()[-1])
for c in d['cases']: print(c)"; tail -3 gpurun_out/bench_local_n2.err | cut -c1-300
