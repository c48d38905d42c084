#!/bin/bash
# round 2, call F: ncu of the frame-lane kernels at full configs[3] size (one launch each of phase A and B)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 1 --warmup 3"
timeout 900 python bench.py $B > gpurun_out/r2f_plain.json 2> gpurun_out/r2f_plain.err &&
timeout 2000 ncu --set full --clock-control none --import-source on -k regex:kf_frames -s 8 -c 2 -o gpurun_out/r2f_kf python bench.py $B > gpurun_out/r2f_ncu.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/r2f_plain.json").read().strip().split("\n")[-1])
print(round(d["value"]), d["ms_per_step"], d["stage_ms"])
PY
tail -3 gpurun_out/r2f_ncu.log
