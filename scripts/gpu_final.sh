mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemv_topk -s 8 -c 1 --csv --log-file gpurun_out/k2_cfg4_traffic.csv \
    python bench.py --workload cfg4 --steps 3 --warmup 5 --no-cpu-baseline --no-extra > /dev/null 2>&1
grep -v "^==" gpurun_out/k2_cfg4_traffic.csv | cut -d, -f5,12- | tail -4
