mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -3
FCS_TC_VERBOSE=1 timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -9 | cut -c1-330
timeout 600 python bench.py --workload cfg3 --rows 1250000 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
