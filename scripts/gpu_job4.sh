mkdir -p gpurun_out
timeout 300 python tests/quick_bench.py upload 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gemv_gpu.py -x -q 2>&1 | tail -2
