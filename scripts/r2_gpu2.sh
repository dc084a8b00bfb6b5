#!/bin/bash
# launch list of one cfg3 search (ncu, serialised) for the share of every kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg3_raw.csv \
   python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_cfg3_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02_launches_cfg3_raw.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ii = hdr.index('ID')
seq = [(r[ki].split('(')[0][-60:], float(r[vi].replace(',', ''))) for r in rows[1:]]
# last search = from the last tc_prep_kernel on
idx = max(i for i, (k, v) in enumerate(seq) if 'tc_prep' in k)
for k, v in seq[idx:idx + 16]:
    print(f'{v/1000:10.1f} us  {k}')
PY
