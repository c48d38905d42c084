#!/bin/bash
# round 2, call AB: the small-batch LPC rule (both channels, four lanes from order 5 up) on larger resident chunks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 20 --warmup 5"
run() {  # name workload scale [ENV=VAL ...]
  local name=$1 w=$2 sc=$3; shift 3
  env "$@" timeout 200 python bench.py $B --workload $w --scale $sc > gpurun_out/r2ab_$name.json 2> gpurun_out/r2ab_$name.err
}
run c2_q5 config2 1 ALACGPU_SMALL_BATCH_FRAMES=100000
run c2_q9 config2 1 ALACGPU_QUAD_MIN_LAST=9 ALACGPU_QUAD_MIN_FIRST=9
run c2_q13 config2 1 ALACGPU_QUAD_MIN_LAST=13 ALACGPU_QUAD_MIN_FIRST=13
run c2_mid config2 1 X=1
for sc in 16 24 30; do
  run c1_s${sc}_q5 config1 $sc ALACGPU_SMALL_BATCH_FRAMES=100000
  run c1_s${sc}_mid config1 $sc X=1
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2ab_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f.split("/")[-1], d["config"]["frames"], round(d["device_ms_per_step"],3))
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json",".err")).read()[-300:])
PY
