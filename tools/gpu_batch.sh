# usage: bash tools/gpu_batch.sh <flags>...  -- the configs[3]-shaped batch leg only, per fusion flag set
for f in "$@"; do
  timeout 300 python bench.py --flags $f --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bb.json 2> gpurun_out/bb.err || tail -3 gpurun_out/bb.err
  python -c "
import json; d=json.load(open('gpurun_out/bb.json'))['batch']; print('batch flags $f: value', round(d['value']), d['stage_ms'])"
done
