mkdir -p gpurun_out
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2973$n bench.py --gpus $n --workload cfg3 --steps 10 --warmup 3 > gpurun_out/scale_cfg3_n${n}_auto.json 2> gpurun_out/scale_cfg3_n${n}_auto.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_cfg3_n${n}_auto.json").read().strip().splitlines()[-1])
    print("n=$n", "value %.0f q/s" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["config"]["parallelism"], "roof %.3f" % d["roofline"]["frac"])
except Exception as e:
    print("n=$n FAILED", e); print(open("gpurun_out/scale_cfg3_n${n}_auto.err").read()[-1500:])
PY
done
