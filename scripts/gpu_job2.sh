mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_gemv_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --workload cfg2 --no-cpu-baseline > gpurun_out/bench_cfg2b.json 2> gpurun_out/bench_cfg2b.err; tail -2 gpurun_out/bench_cfg2b.err; cut -c1-1500 gpurun_out/bench_cfg2b.json
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:tc_gemm_filter -s 24 -c 8 --csv --log-file gpurun_out/k3_rounds_traffic.csv \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
./scripts/micro/umma_two_issuers > gpurun_out/micro_umma_two_issuers.txt 2>&1
./scripts/micro/umma_ld_contention > gpurun_out/micro_umma_ld_contention.txt 2>&1
