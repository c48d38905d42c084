#!/bin/bash
# round 2, call C: per-SM schedule of the frame-lane kernels + threaded staging pipeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2c_tests.log
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 5 --warmup 3"
timeout 600 python bench.py --workload config4 --tracks 64 $B > gpurun_out/r2c_b64_kf.json 2> gpurun_out/r2c_b64_kf.err
ALACGPU_KF_SEGMENTS=74 timeout 600 python bench.py --workload config4 --tracks 64 $B > gpurun_out/r2c_b64_kf74.json 2> gpurun_out/r2c_b64_kf74.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r2c_full.json 2> gpurun_out/r2c_full.err
for f in b64_kf b64_kf74 full; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c_$f.json").read().strip().split("\n")[-1])
    print("$f", round(d["value"]), d["ms_per_step"], d["stage_ms"], (d.get("e2e") or {}).get("value"), (d.get("e2e_pageable") or {}).get("value"))
    if "latency_legs" in d:
        for k,v in d["latency_legs"].items(): print("   ",k, v["device_ms"], v["wall_ms"], v["e2e_ms"])
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2c_$f.err").read()[-1500:])
PY
done
