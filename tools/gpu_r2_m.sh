#!/bin/bash
# round 2, call M: frame lanes forced on the latency configs; per-order cost curve of the frame lanes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 20 --warmup 3"
for w in config1 config2 config3; do
  for fl in 0 128; do
    timeout 300 python bench.py $B --workload $w --flags $fl > gpurun_out/r2m_${w}_$fl.json 2> gpurun_out/r2m_${w}_$fl.err
  done
done
export ALACGPU_KF_MIN=65536
for m in 1 4 8 12 16 20 24 28 30; do
  timeout 300 python bench.py --no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 3 --warmup 3 --workload fixed --orders $m,$m --tracks 128 --unique 4 > gpurun_out/r2m_o$m.json 2> gpurun_out/r2m_o$m.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2m_*.json")):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
        print(f.split("/")[-1], round(d["value"]), round(d["ms_per_step"],3), round(d["device_ms_per_step"],3), {k:round(v,2) for k,v in d["stage_ms"].items()}, d["config"]["decode_path"])
    except Exception as e:
        print(f, "failed", e); print(open(f.replace(".json",".err")).read()[-400:])
PY
