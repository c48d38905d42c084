#!/bin/bash
# round 2, call Y (8 GPUs): configs[4] at full size (10,000 tracks) cut frame-wise over 8 ranks -- final r2 build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2y_box.txt; nproc >> gpurun_out/r2y_box.txt; free -g >> gpurun_out/r2y_box.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 8 --steps 4 --warmup 3 --e2e-steps 1 --e2e-pageable-steps 1 > gpurun_out/r2y_n8.json 2> gpurun_out/r2y_n8.err; echo "n8 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2y_n8.json").read().strip().split("\n")[-1])
    print("n8", round(d["value"]), round(d["ms_per_step"],2), round(d["device_ms_per_step"],2), d["step_wall_ms_rank0"], (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("host_copy_ceiling_gbs_each_way"), (d.get("e2e_pageable") or {}).get("value"), d["config"].get("rank_frame_ranges"), d["config"].get("multi_device_context_check"), d["config"].get("tracks"), d.get("cpu_baseline"))
except Exception as e:
    print("n8 failed", e); print(open("gpurun_out/r2y_n8.err").read()[-2500:])
PY
cat gpurun_out/r2y_box.txt | tail -4
