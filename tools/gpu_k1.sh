# usage: bash tools/gpu_k1.sh <tag>   -- parity tests, A/B bench (three kernels vs fused), source-level ncu of k1_entropy
tag=$1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in 2 0; do
  timeout 200 python bench.py --flags $f --steps 10 --warmup 3 --no-cpu --e2e-steps 3 --batch-tracks 0 > gpurun_out/p$f.json 2> gpurun_out/p$f.err || tail -3 gpurun_out/p$f.err
  python -c "
import json; d=json.load(open('gpurun_out/p$f.json')); print('run', $f, round(d['value']), d['stage_ms'], round(d['e2e']['value']))"
done
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section LaunchStats --section SpeedOfLight --clock-control none --import-source on -k regex:k1_entropy --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_${tag}_k1 python bench.py --flags 2 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --batch-tracks 0 > gpurun_out/ncu_${tag}_k1.log 2>&1
tail -1 gpurun_out/ncu_${tag}_k1.log
