#!/bin/bash
# round 2, call G: regression tests, the full default bench line, pageable copy-thread variants, reference arm, ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2g_tests.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err
P="--no-cpu --e2e-steps 0 --e2e-pageable-steps 2 --latency-steps 0 --steps 1 --warmup 3"
for t in 6 14; do ALACGPU_COPY_THREADS=$t timeout 600 python bench.py $P > gpurun_out/r2g_pg$t.json 2> gpurun_out/r2g_pg$t.err; done
python - <<PY
import json
for f in ("bench","pg6","pg14","ref"):
    try:
        d=json.loads(open(f"gpurun_out/r2g_{f}.json").read().strip().split("\n")[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), (d.get("e2e") or {}).get("value"), (d.get("e2e_pageable") or {}).get("value"), d.get("cpu_baseline",{}).get("value"))
        for k,v in (d.get("latency_legs") or {}).items(): print("   ",k, round(v["device_ms"],3), round(v["wall_ms"],3), round(v["e2e_ms"],3))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2g_{f}.err").read()[-600:])
PY
N="--workload config4 --tracks 250 --no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 1 --warmup 3"
export ALACGPU_KF_MIN=65536
timeout 600 python bench.py $N > gpurun_out/r2g_n250.json 2> gpurun_out/r2g_n250.err &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:kf_frames -s 8 -c 2 -o gpurun_out/r2g_kf250 python bench.py $N > gpurun_out/r2g_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv python bench.py $N > gpurun_out/r2g_ncu2.log 2>&1
tail -2 gpurun_out/r2g_ncu.log; ls -la gpurun_out/r2g_*
