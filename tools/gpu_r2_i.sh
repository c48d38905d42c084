#!/bin/bash
# round 2, call I: validation of the short channel-A ring / six phase-B blocks, copy ceiling, full default line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_frame_lanes.py tests/test_gpu_checked.py -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2i_tests.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 4 --warmup 3"
ALACGPU_KF_MIN=65536 timeout 600 python bench.py --tracks 250 $B > gpurun_out/r2i_t250.json 2> gpurun_out/r2i_t250.err
python - <<PY
import json
for f in ("bench","t250"):
    try:
        d=json.loads(open(f"gpurun_out/r2i_{f}.json").read().strip().split("\n")[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["device_ms_per_step"],2), {k:round(v,1) for k,v in d["stage_ms"].items()}, (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("host_copy_ceiling_gbs_each_way"), (d.get("e2e") or {}).get("achieved_copy_gbs_each_way"), (d.get("e2e_pageable") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2i_{f}.err").read()[-800:])
PY
