mkdir -p gpurun_out
SECONDS=0
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_default2.json 2> gpurun_out/bench_default2.err
echo "exit $? in ${SECONDS}s"; tail -2 gpurun_out/bench_default2.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default2.json"))
print("primary", d["config"]["workload"], "%.0f q/s" % d["value"], "frac %.3f" % d["roofline"]["frac"], "e2e %.0f" % d["e2e"]["value"])
for k,v in d["extra"].items():
    print(" extra", k, ("%.1f q/s ms %.3f frac %.3f e2e %.1f fb %s" % (v["value"], v["ms_per_step"], v["roofline"]["frac"], v["e2e"]["value"], v["config"]["tc_fallback_queries"])) if "value" in v else v)
PY
