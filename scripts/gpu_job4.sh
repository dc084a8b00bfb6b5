mkdir -p gpurun_out
timeout 120 python bench.py --workload cfg3 --rows 1250000 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -4 | cut -c1-600
echo "exit: $?"
