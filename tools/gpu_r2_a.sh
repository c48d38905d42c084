#!/bin/bash
# round 2, call A: frame-lane parity + first throughput numbers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_box.txt 2>&1
nproc >> gpurun_out/r2a_box.txt; free -g >> gpurun_out/r2a_box.txt
timeout 900 python -m pytest tests/test_gpu_frame_lanes.py -x -q > gpurun_out/r2a_tests_fl.log 2>&1; echo "fl tests rc=$?" >> gpurun_out/r2a_box.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_host_mirror.py -x -q -m gpu > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_box.txt
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 5 --warmup 3"
timeout 600 python bench.py --workload config4 --tracks 64 $B --flags 64 > gpurun_out/r2a_b64_old.json 2> gpurun_out/r2a_b64_old.err
timeout 600 python bench.py --workload config4 --tracks 64 $B > gpurun_out/r2a_b64_kf.json 2> gpurun_out/r2a_b64_kf.err
timeout 600 python bench.py --workload config4 --tracks 200 $B > gpurun_out/r2a_b200_kf.json 2> gpurun_out/r2a_b200_kf.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_full.json 2> gpurun_out/r2a_full.err
tail -c 600 gpurun_out/r2a_box.txt; tail -3 gpurun_out/r2a_tests_fl.log; tail -3 gpurun_out/r2a_tests.log
for f in b64_old b64_kf b200_kf full; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2a_$f.json").read().strip().split("\n")[-1])
    print("$f", round(d["value"]), d["ms_per_step"], d["stage_ms"], (d.get("e2e") or {}).get("value"), (d.get("e2e_pageable") or {}).get("value"))
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r2a_$f.err").read()[-1500:])
PY
done
