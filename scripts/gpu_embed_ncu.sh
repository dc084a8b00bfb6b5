mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:embed_edge_tc -s 6 -c 1 -o gpurun_out/prof_embed_edge_tc \
    python bench.py --workload embed --nq 512 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_embed_tc.log 2>&1
tail -2 gpurun_out/ncu_embed_tc.log
