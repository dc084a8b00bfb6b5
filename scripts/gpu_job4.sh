mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit $? in ${SECONDS}s"; tail -3 gpurun_out/bench_default.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default.json"))
print("primary", d["config"]["workload"], "%.0f q/s" % d["value"], "frac %.3f" % d["roofline"]["frac"], "e2e %.0f" % d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"])
for k,v in d["extra"].items():
    print(" extra", k, ("%.1f q/s frac %.3f e2e %.1f" % (v["value"], v["roofline"]["frac"], v["e2e"]["value"])) if "value" in v else v)
PY
