#!/bin/bash
# Round-end evidence: full GPU test suite, smoke, default bench line (with extras), cfg5 slice, embed ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default3.json 2> gpurun_out/bench_default3.err; tail -2 gpurun_out/bench_default3.err
timeout 300 python bench.py --workload embed > gpurun_out/bench_embed.json 2> gpurun_out/bench_embed.err; tail -2 gpurun_out/bench_embed.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_embed.csv \
    python bench.py --workload embed --nq 512 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:embed_edge_tc -s 6 -c 1 -o gpurun_out/prof_embed_edge_tc \
    python bench.py --workload embed --nq 512 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_embed.log 2>&1
FCS_EMBED_MODE=0 timeout 300 python bench.py --workload embed --no-cpu-baseline > gpurun_out/bench_embed_fp32.json 2>/dev/null
python - <<'PY'
import json
for f in ("bench_default3", "bench_embed", "bench_embed_fp32"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, d["config"]["workload"], "%.1f %s" % (d["value"], d["unit"]), "ms/step %.3f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
          "e2e %.1f" % d["e2e"]["value"], "fb", d["config"].get("tc_fallback_queries"))
    for k, o in d.get("extra", {}).items():
        if "error" in o: print("  extra", k, "ERROR", o["error"]); continue
        print("  extra", k, "%.1f %s" % (o["value"], o["unit"]), "ms/step %.3f" % o["ms_per_step"], "frac %.3f" % o["roofline"]["frac"], "e2e %.1f" % o["e2e"]["value"])
PY
