#!/bin/bash
# round 2, call R: eight-lane LPC + small-batch policy -- parity suite, latency configs, size sweep for the policy thresholds
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2r_tests.log
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 20 --warmup 5"
run() {  # name workload scale [ENV=VAL ...]
  local name=$1 w=$2 sc=$3; shift 3
  env "$@" timeout 300 python bench.py $B --workload $w --scale $sc > gpurun_out/r2r_$name.json 2> gpurun_out/r2r_$name.err
}
for w in config1 config2 config3; do run def_$w $w 1 X=1; done

# policy alternatives on the small configs
for w in config1 config3; do
  run ${w}_w9 $w 1 ALACGPU_LPC_WIDE=1 ALACGPU_QUAD_MIN_LAST=9 ALACGPU_QUAD_MIN_FIRST=9
  run ${w}_w5 $w 1 ALACGPU_LPC_WIDE=1 ALACGPU_QUAD_MIN_LAST=5 ALACGPU_QUAD_MIN_FIRST=5
  run ${w}_w13 $w 1 ALACGPU_LPC_WIDE=1 ALACGPU_QUAD_MIN_LAST=13 ALACGPU_QUAD_MIN_FIRST=13
  run ${w}_q9 $w 1 ALACGPU_LPC_WIDE=0 ALACGPU_QUAD_MIN_LAST=9 ALACGPU_QUAD_MIN_FIRST=9
  run ${w}_q17 $w 1 ALACGPU_LPC_WIDE=0 ALACGPU_QUAD_MIN_LAST=17 ALACGPU_QUAD_MIN_FIRST=17
done
# size sweep (config1 = 646 frames at scale 1): wide / both-channel quads / last-channel quads only
for sc in 2 4 8 16; do
  run s${sc}_wide config1 $sc ALACGPU_WIDE_MAX_FRAMES=100000
  run s${sc}_both config1 $sc ALACGPU_WIDE_MAX_FRAMES=0 ALACGPU_BOTH_MAX_FRAMES=100000
  run s${sc}_last config1 $sc ALACGPU_WIDE_MAX_FRAMES=0 ALACGPU_BOTH_MAX_FRAMES=0
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2r_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f.split("/")[-1], d["config"]["frames"], round(d["ms_per_step"],3), round(d["device_ms_per_step"],3), {k:round(v,2) for k,v in d["stage_ms"].items()})
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json",".err")).read()[-300:])
PY
