#!/bin/bash
# round 2, call B: occupancy + ncu of the frame-lane kernels on a 32-track batch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--workload config4 --tracks 32 --no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 2 --warmup 3"
ALACGPU_DEBUG_OCC=1 timeout 600 python bench.py $B > gpurun_out/r2b_plain.json 2> gpurun_out/r2b_plain.err &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:kf_frames -s 4 -c 2 -o gpurun_out/r2b_kf python bench.py $B > gpurun_out/r2b_ncu.log 2>&1
tail -5 gpurun_out/r2b_plain.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2b_plain.json").read().strip().split("\n")[-1])
print(round(d["value"]), d["ms_per_step"], d["stage_ms"])
PY
tail -5 gpurun_out/r2b_ncu.log; ls -la gpurun_out/r2b_kf*
