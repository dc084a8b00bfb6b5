mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --workload cfg3 --steps 5 --warmup 3 2>&1 | tail -3 | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --workload cfg2 2>&1 | tail -2 | cut -c1-1200
