#!/bin/bash
# round 2, call P: warp-voted fast entropy step on the stream-lane path -- parity suite, then the latency configs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2p_tests.log
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 20 --warmup 5"
for w in config1 config2 config3; do
  timeout 300 python bench.py $B --workload $w > gpurun_out/r2p_$w.json 2> gpurun_out/r2p_$w.err
  timeout 300 python bench.py $B --workload $w --flags 2 > gpurun_out/r2p_${w}_nofuse.json 2> gpurun_out/r2p_${w}_nofuse.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p_config*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"],3), round(d["device_ms_per_step"],3), {k:round(v,2) for k,v in d["stage_ms"].items()}, d["config"]["decode_path"])
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json",".err")).read()[-400:])
PY
