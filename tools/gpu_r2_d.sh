#!/bin/bash
# round 2, call D: ncu of the per-SM-scheduled frame-lane kernels (32 tracks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--workload config4 --tracks 32 --no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 2 --warmup 3"
timeout 600 python bench.py $B > gpurun_out/r2d_plain.json 2> gpurun_out/r2d_plain.err &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:kf_frames -s 4 -c 2 -o gpurun_out/r2d_kf python bench.py $B > gpurun_out/r2d_ncu.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/r2d_plain.json").read().strip().split("\n")[-1])
print(round(d["value"]), d["ms_per_step"], d["stage_ms"])
PY
tail -3 gpurun_out/r2d_ncu.log
