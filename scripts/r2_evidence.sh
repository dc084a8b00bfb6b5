#!/bin/bash
# Round 2 evidence (1 GPU): ncu launch lists, per-launch DRAM traffic + tensor-pipe activity of the K3 rounds, --set full of the sweep
mkdir -p gpurun_out
B="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-extra"
echo "== launch list cfg3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg3_raw.csv $B > gpurun_out/ncu_a.log 2>&1
echo "== K3 rounds: dram bytes + tensor pipe"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:tc_gemm_filter -s 15 -c 5 --csv --log-file gpurun_out/r02_k3_rounds_raw.csv $B > gpurun_out/ncu_b.log 2>&1
echo "== --set full of the sweep"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_filter -s 19 -c 1 -f -o gpurun_out/r02_prof_k3_sweep $B > gpurun_out/ncu_c.log 2>&1
tail -2 gpurun_out/ncu_c.log
echo "== launch list + traffic cfg2 / cfg4"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemv_topk -s 20 -c 3 --csv --log-file gpurun_out/r02_k2_cfg2_raw.csv python bench.py --workload cfg2 --steps 20 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/ncu_d.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemv_topk -s 12 -c 2 --csv --log-file gpurun_out/r02_k2_cfg4_raw.csv python bench.py --workload cfg4 --steps 5 --warmup 10 --no-cpu-baseline --no-extra > gpurun_out/ncu_e.log 2>&1
ls -la gpurun_out/*.csv gpurun_out/*.ncu-rep | tail
