#!/bin/bash
# round 2, call N: final evidence with the final build -- tests, default bench line, reference arm, ncu at 250 tracks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2n_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -2 gpurun_out/r2n_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_ref.json 2> gpurun_out/r2n_ref.err
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0"
ALACGPU_KF_MIN=65536 timeout 300 python bench.py $B --steps 10 --warmup 3 --workload fixed --orders 20,20 --tracks 128 --unique 4 > gpurun_out/r2n_o20.json 2> gpurun_out/r2n_o20.err
python - <<PY
import json
for f in ("bench","ref","o20"):
    try:
        d=json.loads(open(f"gpurun_out/r2n_{f}.json").read().strip().split("\n")[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), d.get("device_ms_per_step"), d.get("step_wall_ms_rank0"), (d.get("e2e") or {}).get("value"), (d.get("e2e_pageable") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
        for k,v in (d.get("latency_legs") or {}).items(): print("   ",k, round(v["device_ms"],3), round(v["wall_ms"],3), round(v["e2e_ms"],3))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2n_{f}.err").read()[-600:])
PY
N="--workload config4 --tracks 250 $B --steps 1 --warmup 3"
export ALACGPU_KF_MIN=65536
timeout 600 python bench.py $N > gpurun_out/r2n_n250.json 2> gpurun_out/r2n_n250.err &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:kf_frames -s 8 -c 2 -o gpurun_out/r2n_kf250 python bench.py $N > gpurun_out/r2n_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py $N > gpurun_out/r2n_ncu2.log 2>&1
tail -2 gpurun_out/r2n_ncu.log
