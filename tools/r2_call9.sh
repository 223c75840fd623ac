#!/bin/bash
# Round-2 GPU call 9: bench lines of every BASELINE.json configuration at N = 1
set -u
mkdir -p gpurun_out
for c in headline c2 c3 c4 c5 c1; do
  steps=24; [ $c = c2 ] && steps=12; [ $c = c4 ] && steps=5; [ $c = c1 ] && steps=50
  timeout 900 python bench.py --config $c --steps $steps --warmup 3 > gpurun_out/r2c9_bench_$c.json 2> gpurun_out/r2c9_bench_$c.err
  echo "== $c rc=$?"; tail -c 600 gpurun_out/r2c9_bench_$c.err; python - $c <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2c9_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(d["metric"], d["value"], d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", d["e2e"] and d["e2e"]["value"],
          "| roofline", r["kernel"][:16], round(r["frac"], 3), "ms", round(r["ms_per_launch"], 3), "share", round(r["share_of_step"], 3),
          "| cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "| fp/bp gproj", d.get("fp_gproj_per_s"), d.get("bp_gproj_per_s"))
except Exception as e:
    print("no line", e)
PY
done
