#!/bin/bash
# round 2, call Q: latency configs -- four-lane LPC thresholds (last / first channel), old vs new entropy step, unfused stage times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 20 --warmup 5"
run() {  # name workload last first [extra...]
  local name=$1 w=$2 l=$3 f=$4; shift 4
  ALACGPU_QUAD_MIN_LAST=$l ALACGPU_QUAD_MIN_FIRST=$f timeout 300 python bench.py $B --workload $w "$@" > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err
}
for lf in "17 0" "17 17" "13 13" "9 9" "5 5" "25 25" "9 17" "5 13"; do set -- $lf; run c1_l$1_f$2 config1 $1 $2; done
for l in 17 13 9 5 25; do run c3_l$l config3 $l 0; done
for lf in "17 0" "25 0" "17 25"; do set -- $lf; run c2_l$1_f$2 config2 $1 $2; done
export ALACGPU_LIB=$PWD/alac/net_b200/libalacgpu_oldstep.so
for w in config1 config2 config3; do
  run old_$w $w 17 0
  run old_${w}_nofuse $w 17 0 --flags 2
done
unset ALACGPU_LIB
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2q_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f.split("/")[-1], round(d["ms_per_step"],3), round(d["device_ms_per_step"],3), {k:round(v,2) for k,v in d["stage_ms"].items()})
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json",".err")).read()[-300:])
PY
