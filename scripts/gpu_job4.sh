mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -2
for nq in 512 4096; do
FCS_TC_VERBOSE=0 timeout 600 python bench.py --workload cfg3 --nq $nq --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('nq $nq: %.3f ms/step  %.0f q/s  K3 frac %.3f  launches %d fallbacks %d' % (d['ms_per_step'], d['value'], d['roofline']['frac'], d['gpu_launches'], d['config']['tc_fallback_queries']))"
done
