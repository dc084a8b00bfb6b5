mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_embed_gpu.py -x -q -s 2>&1 | tail -25
