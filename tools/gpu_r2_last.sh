#!/bin/bash
# round 2, very last call: the GPU suite and the default bench line with the committed sources
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2l_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2l_bench.json").read().strip().split("\n")[-1])
print(round(d["value"]), round(d["ms_per_step"],2), (d.get("e2e") or {}).get("value"), (d.get("e2e_pageable") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
for k,v in (d.get("latency_legs") or {}).items(): print("   ",k, round(v["device_ms"],3), round(v["e2e_ms"],3))
PY
