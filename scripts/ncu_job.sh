#!/bin/bash
# ncu evidence for profiles/: launch lists (device time per launch) and one full capture per dominant kernel.
mkdir -p gpurun_out
set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg3_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches_cfg2.csv \
    python bench.py --workload cfg2 --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_cfg2_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_filter -s 30 -c 1 -o gpurun_out/prof_k3 \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_topk -s 15 -c 1 -o gpurun_out/prof_k2 \
    python bench.py --workload cfg2 --steps 10 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_k2.log 2>&1
ls -la gpurun_out
