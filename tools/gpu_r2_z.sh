#!/bin/bash
# round 2, call Z: value-only round in the frame-lane step -- parity suite, then configs[3] with and without it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2z_tests.log
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 8 --warmup 3"
timeout 400 python bench.py $B > gpurun_out/r2z_fast.json 2> gpurun_out/r2z_fast.err; echo "fast rc=$?"
ALACGPU_LIB=$PWD/alac/net_b200/libalacgpu_nofast.so timeout 400 python bench.py $B > gpurun_out/r2z_nofast.json 2> gpurun_out/r2z_nofast.err; echo "nofast rc=$?"
python - <<PY
import json
for f in ("fast","nofast"):
    try:
        d=json.loads(open(f"gpurun_out/r2z_{f}.json").read().strip().split("\n")[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()})
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2z_{f}.err").read()[-600:])
PY
