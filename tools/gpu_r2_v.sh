#!/bin/bash
# round 2, call V: per-warp start / end trace of the fused launch on configs[1]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 1 --warmup 1"
ALACGPU_TRACE=gpurun_out/r2v_trace.txt timeout 300 python bench.py $B --workload config2 > gpurun_out/r2v.json 2> gpurun_out/r2v.err
python tools/trace_summary.py gpurun_out/r2v_trace.txt all | head -60
