mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -8
