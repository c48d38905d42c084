#!/bin/bash
# round 2, call E: checked build, whole-shard chunks, occupancy variants, old path at full size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_checked.py tests/test_gpu_frame_lanes.py -x -q -m gpu > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2e_tests.log
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 4 --warmup 3"
run() { name=$1; shift; env "$@" timeout 900 python bench.py $B $EXTRA > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2e_$name.json").read().strip().split("\n")[-1])
    print("$name", round(d["value"]), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms"].items()}, d["config"]["chunks_per_step"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2e_$name.err").read()[-800:])
PY
}
EXTRA="" run full_kf A=1
EXTRA="--flags 64" run full_old A=1
EXTRA="--tracks 250" run t250_kf A=1
EXTRA="--tracks 250" run t250_padA9 ALACGPU_KF_PAD_A=9
EXTRA="--tracks 250" run t250_padA26 ALACGPU_KF_PAD_A=26
EXTRA="--tracks 250" run t250_padB14 ALACGPU_KF_PAD_B=14
EXTRA="--tracks 250" run t250_padA9B14 ALACGPU_KF_PAD_A=9 ALACGPU_KF_PAD_B=14
EXTRA="--tracks 250 --flags 64" run t250_old A=1
