#!/bin/bash
# round 2, call AA: the small-batch LPC rule on 24-bit material and around its frame threshold
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 20 --warmup 5"
run() {  # name workload scale [ENV=VAL ...]
  local name=$1 w=$2 sc=$3; shift 3
  env "$@" timeout 200 python bench.py $B --workload $w --scale $sc > gpurun_out/r2aa_$name.json 2> gpurun_out/r2aa_$name.err
}
for sc in 0.1 0.25; do
  run c2_s${sc}_small config2 $sc X=1
  run c2_s${sc}_mid config2 $sc ALACGPU_SMALL_BATCH_FRAMES=0
done
for sc in 8 12; do
  run c1_s${sc}_small config1 $sc ALACGPU_SMALL_BATCH_FRAMES=100000
  run c1_s${sc}_mid config1 $sc ALACGPU_SMALL_BATCH_FRAMES=0
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2aa_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f.split("/")[-1], d["config"]["frames"], round(d["device_ms_per_step"],3))
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json",".err")).read()[-300:])
PY
