#!/bin/bash
# round 2, call T: final evidence with the final build -- tests, smoke, default bench line (configs[3] + latency legs), reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2t_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2t_smoke.log 2>&1; tail -2 gpurun_out/r2t_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2t_ref.json 2> gpurun_out/r2t_ref.err
python - <<PY
import json
for f in ("bench","ref"):
    try:
        d=json.loads(open(f"gpurun_out/r2t_{f}.json").read().strip().split("\n")[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), d.get("device_ms_per_step"), d.get("step_wall_ms_rank0"), (d.get("e2e") or {}).get("value"), (d.get("e2e_pageable") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
        for k,v in (d.get("latency_legs") or {}).items(): print("   ",k, round(v["device_ms"],3), round(v["wall_ms"],3), round(v["e2e_ms"],3))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2t_{f}.err").read()[-600:])
PY
