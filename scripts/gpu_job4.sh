mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py -x -q 2>&1 | tail -2
FCS_TC_VERBOSE=1 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | tail -10 | cut -c1-300
