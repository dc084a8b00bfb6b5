#!/bin/bash
# First GPU call of the next round (1 GPU): what round 1 could not measure any more, and the cheap sweeps DESIGN.md §9 asks for.
#   /usr/local/graft/bin/gpurun --timeout 900 -- bash scripts/r2_first_call.sh
mkdir -p gpurun_out
# 1. full suite on the final round-1 tree
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
# 2. growth factor per batch size (FCS_TC_GROWTH was only swept at 4096 queries; N=8 strong scaling runs 512 per rank)
for nq in 512 1024 4096; do for g in 1.5 2 3 4 6; do
  FCS_TC_GROWTH=$g timeout 300 python bench.py --workload cfg3 --nq $nq --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('nq $nq growth $g: %.3f ms/step  %.0f q/s  K3 frac %.3f  fallbacks %d' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['config']['tc_fallback_queries']))"
done; done
# 3. ncu of the final embedder revision (the committed capture is one revision older)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:embed_edge_tc -s 6 -c 1 -o gpurun_out/prof_embed_edge_tc_r2 \
    python bench.py --workload embed --nq 512 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_embed_r2.log 2>&1
tail -2 gpurun_out/ncu_embed_r2.log
# (2 GPUs, separate call: --gpus 2)  python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 scripts/dist_embed_check.py
# (8 GPUs, separate call: --gpus 8)  torchrun ... bench.py --gpus 8 --workload cfg5 --no-cpu-baseline --no-extra
