#!/bin/bash
# round 2, call W: dealt LPC blocks on mid-size resident chunks -- env-case parity, configs[1] with / without, trace, size sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tuning_environment or config2 or full_size" > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2w_tests.log
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 20 --warmup 5"
run() {  # name workload scale [ENV=VAL ...]
  local name=$1 w=$2 sc=$3; shift 3
  env "$@" timeout 300 python bench.py $B --workload $w --scale $sc > gpurun_out/r2w_$name.json 2> gpurun_out/r2w_$name.err
}
run c2_deal config2 1 X=1
run c2_nodeal config2 1 ALACGPU_LPC_DEAL=0
run c2_deal_f25 config2 1 ALACGPU_QUAD_MIN_FIRST=25
run c2_deal_f17 config2 1 ALACGPU_QUAD_MIN_FIRST=17
run c2_deal_l13 config2 1 ALACGPU_QUAD_MIN_LAST=13
run c2_deal_noquad config2 1 ALACGPU_QUAD_MIN_LAST=0 ALACGPU_QUAD_MIN_FIRST=0
B2="$B --flags 2"; env X=1 timeout 300 python bench.py $B2 --workload config2 > gpurun_out/r2w_c2_deal_nofuse.json 2> gpurun_out/r2w_c2_deal_nofuse.err
for sc in 8 16 24; do
  run s${sc}_deal config1 $sc X=1
  run s${sc}_nodeal config1 $sc ALACGPU_LPC_DEAL=0
done
ALACGPU_TRACE=gpurun_out/r2w_trace.txt timeout 300 python bench.py --no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 1 --warmup 1 --workload config2 > gpurun_out/r2w_tr.json 2> gpurun_out/r2w_tr.err
python tools/trace_summary.py gpurun_out/r2w_trace.txt | head -8
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2w_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f.split("/")[-1], d["config"]["frames"], round(d["device_ms_per_step"],3), {k:round(v,2) for k,v in d["stage_ms"].items()})
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json",".err")).read()[-300:])
PY
