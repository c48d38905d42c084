#!/bin/bash
# round 2, call S: small-batch LPC policy -- lanes per stream (4 / 8) x smallest multi-lane order (5 / 9) over batch sizes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 20 --warmup 5"
run() {  # name workload scale [ENV=VAL ...]
  local name=$1 w=$2 sc=$3; shift 3
  env "$@" timeout 300 python bench.py $B --workload $w --scale $sc > gpurun_out/r2s_$name.json 2> gpurun_out/r2s_$name.err
}
for cfg in "config1 1" "config1 2" "config1 4" "config1 6" "config3 1" "config3 2" "config3 4" "config3 8"; do
  set -- $cfg
  for pol in "q 0 5" "q 0 9" "w 1 5" "w 1 9" "q 0 13"; do
    set -- $cfg $pol
    run $1_s$2_$3$5 $1 $2 ALACGPU_LPC_WIDE=$4 ALACGPU_QUAD_MIN_LAST=$5 ALACGPU_QUAD_MIN_FIRST=$5
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2s_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f.split("/")[-1], d["config"]["frames"], round(d["device_ms_per_step"],3), round(d["stage_ms"]["entropy_ms"],3))
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json",".err")).read()[-300:])
PY
