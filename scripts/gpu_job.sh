set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 300 python tests/quick_bench.py 2>&1 | tail -12
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -3 gpurun_out/bench_cfg3.err; cat gpurun_out/bench_cfg3.json
timeout 600 python bench.py --workload cfg2 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -3 gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json
