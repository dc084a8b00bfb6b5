mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 --workload cfg3 --steps 10 --warmup 3 > gpurun_out/scale_cfg3_n8.json 2> gpurun_out/scale_cfg3_n8.err
echo "exit $?"
tail -1 gpurun_out/scale_cfg3_n8.json | cut -c1-400
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/scale_cfg3_n8.err | tail -25 | cut -c1-300
