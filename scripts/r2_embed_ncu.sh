#!/bin/bash
# ncu --set full of the default tensor-core edge kernel of the query embedder (embed_edge_tc2_kernel<16>)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:embed_edge_tc2 -s 6 -c 1 -f -o gpurun_out/r02_prof_embed_edge_tc2 \
    python bench.py --workload embed --nq 512 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_embed_r2.log 2>&1
tail -2 gpurun_out/ncu_embed_r2.log
ls -la gpurun_out/r02_prof_embed_edge_tc2.ncu-rep
