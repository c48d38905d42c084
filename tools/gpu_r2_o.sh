#!/bin/bash
# round 2, call H (2 GPUs): multi-device context tests, configs[4] frame-sharded over 2 ranks, single-process mode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2o_box.txt; nproc >> gpurun_out/r2o_box.txt; free -g >> gpurun_out/r2o_box.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_devices" > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2o_tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 2 --e2e-pageable-steps 1 > gpurun_out/r2o_n2.json 2> gpurun_out/r2o_n2.err; echo "n2 rc=$?"
timeout 1500 python bench.py --gpus 2 --single-process --steps 3 --warmup 3 --no-cpu --latency-steps 0 --e2e-steps 2 --e2e-pageable-steps 0 > gpurun_out/r2o_sp2.json 2> gpurun_out/r2o_sp2.err; echo "sp2 rc=$?"
python - <<PY
import json
for f in ("n2","sp2"):
    try:
        d=json.loads(open(f"gpurun_out/r2o_{f}.json").read().strip().split("\n")[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],2), (d.get("e2e") or {}).get("value"), (d.get("e2e_pageable") or {}).get("value"), d["config"].get("rank_frame_ranges"), d["config"].get("multi_device_context_check"), d["config"].get("tracks"))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2o_{f}.err").read()[-1500:])
PY
