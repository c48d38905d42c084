#!/bin/bash
# round 2, call K: validation after the residual-counter merge; per-step wall times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2k_tests.log
timeout 1200 python bench.py --steps 10 --warmup 3 --e2e-steps 2 --e2e-pageable-steps 1 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2k_bench.json").read().strip().split("\n")[-1])
print(round(d["value"]), round(d["ms_per_step"],2), round(d["device_ms_per_step"],2), {k:round(v,1) for k,v in d["stage_ms"].items()})
print(d["step_wall_ms_rank0"]); print(d["decode_all_ms_rank0"])
print((d.get("e2e") or {}).get("value"), (d.get("e2e_pageable") or {}).get("value"))
PY
