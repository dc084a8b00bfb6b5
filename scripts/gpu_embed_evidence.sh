#!/bin/bash
# Embedder evidence: bench line, ncu launch list, one full capture of embed_edge_kernel.
mkdir -p gpurun_out
timeout 300 python bench.py --workload embed > gpurun_out/bench_embed.json 2> gpurun_out/bench_embed.err; tail -3 gpurun_out/bench_embed.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_embed.csv \
    python bench.py --workload embed --nq 512 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:embed_edge -s 6 -c 1 -o gpurun_out/prof_embed_edge \
    python bench.py --workload embed --nq 512 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_embed.log 2>&1
tail -3 gpurun_out/ncu_embed.log
cut -c1-1800 gpurun_out/bench_embed.json
