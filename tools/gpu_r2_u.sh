#!/bin/bash
# round 2, call U: ncu --set full of the fused entropy + LPC launch on the three latency configs (final build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--no-cpu --e2e-steps 0 --e2e-pageable-steps 0 --latency-steps 0 --steps 2 --warmup 3"
for w in config1 config2 config3; do
  timeout 300 python bench.py $B --workload $w > gpurun_out/r2u_$w.json 2> gpurun_out/r2u_$w.err &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k12_entropy_lpc -s 4 -c 1 -f -o gpurun_out/r2u_k12_$w python bench.py $B --workload $w > gpurun_out/r2u_ncu_$w.log 2>&1
  tail -1 gpurun_out/r2u_ncu_$w.log
done
ls -la gpurun_out/r2u_k12_*.ncu-rep
