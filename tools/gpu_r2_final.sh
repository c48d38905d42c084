#!/bin/bash
# round 2, last call: smoke() and the default bench line with the committed sources
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2f_bench.json").read().strip().split("\n")[-1])
print(round(d["value"]), round(d["ms_per_step"],2), (d.get("e2e") or {}).get("value"), (d.get("e2e_pageable") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
for k,v in (d.get("latency_legs") or {}).items(): print("   ",k, round(v["device_ms"],3), round(v["e2e_ms"],3), (v.get("issue") or {}).get("issue_frac"))
PY
