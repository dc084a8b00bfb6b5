mkdir -p gpurun_out
FCS_TEST_EMBED_TC=1 timeout 150 python -m pytest tests/test_embed_tc_gpu.py -x -q -s 2>&1 | tail -25
echo "exit: $?"
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
