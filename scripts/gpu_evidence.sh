#!/bin/bash
# Round evidence: gpu tests, smoke, bench lines for the three workloads, ncu launch lists + full captures.
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py --workload cfg3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -2 gpurun_out/bench_cfg3.err
timeout 600 python bench.py --workload cfg2 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -2 gpurun_out/bench_cfg2.err
timeout 600 python bench.py --workload cfg4 --steps 50 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; tail -2 gpurun_out/bench_cfg4.err
timeout 600 python bench.py --impl reference --workload cfg3 --steps 1 > gpurun_out/bench_ref_cfg3.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_cfg2.csv \
    python bench.py --workload cfg2 --steps 40 --warmup 10 --no-cpu-baseline --no-extra > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_filter -s 31 -c 1 -o gpurun_out/prof_k3 \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_topk -s 15 -c 1 -o gpurun_out/prof_k2 \
    python bench.py --workload cfg2 --steps 10 --warmup 10 --no-cpu-baseline --no-extra > /dev/null 2>&1
cat gpurun_out/bench_cfg3.json gpurun_out/bench_cfg2.json gpurun_out/bench_cfg4.json gpurun_out/bench_ref_cfg3.json | cut -c1-2500
