mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
SECONDS=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $? in ${SECONDS}s"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --impl reference --steps 1 2>/dev/null | tail -1 | cut -c1-200
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print("primary", d["config"]["workload"], "%.0f q/s" % d["value"], "frac %.3f" % d["roofline"]["frac"], "e2e %.0f" % d["e2e"]["value"], d["config"]["parallelism"])
    for k,v in d["extra"].items():
        print(" extra", k, ("%.1f q/s frac %.3f e2e %.1f %s" % (v["value"], v["roofline"]["frac"], v["e2e"]["value"], v["config"]["parallelism"])) if "value" in v else v)
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_n$N.err").read()[-2000:])
PY
