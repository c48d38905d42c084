#!/bin/bash
# round 2, call L: early-exit predictor groups + quantiser bit in the class key
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_frame_lanes.py tests/test_gpu_checked.py -x -q -m gpu > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2l_tests.log
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 5 --warmup 3"
timeout 900 python bench.py $B > gpurun_out/r2l_full.json 2> gpurun_out/r2l_full.err
timeout 900 python bench.py $B --workload fixed --orders 8,8 --tracks 250 > gpurun_out/r2l_o8.json 2> gpurun_out/r2l_o8.err
python - <<PY
import json
for f in ("full","o8"):
    try:
        d=json.loads(open(f"gpurun_out/r2l_{f}.json").read().strip().split("\n")[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), round(d["device_ms_per_step"],2), {k:round(v,1) for k,v in d["stage_ms"].items()})
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2l_{f}.err").read()[-800:])
PY
