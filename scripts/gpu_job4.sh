mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemv_gpu.py -x -q -k "back_to_back" 2>&1 | tail -15
