# usage: bash tools/gpu_batch.sh <flags>...  -- configs[1] value + the configs[3]-shaped batch leg, per fusion flag set
for f in "$@"; do
  timeout 300 python bench.py --flags $f --steps 5 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/bb.json 2> gpurun_out/bb.err || tail -3 gpurun_out/bb.err
  python -c "
import json; j=json.load(open('gpurun_out/bb.json')); d=j['batch']; print('flags $f: config2 value', round(j['value']), 'k12 %.3f'%j['stage_ms']['entropy_ms'], 'e2e', round(j['e2e']['value']), '| batch value', round(d['value']), 'kernels %.2f ms'%d['stage_ms']['kernels_ms'])"
done
