mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemv_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tests/quick_bench.py 2>&1 | tail -10
timeout 300 python bench.py --workload cfg2 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
