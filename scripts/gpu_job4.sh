mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_gemv_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -6
