# usage: bash tools/gpu_quads.sh "<last first> <last first> ..."  -- A/B of the four-lane LPC thresholds
for pair in "$@"; do
  set -- $pair
  for f in 2 0; do
    ALACGPU_QUAD_MIN_LAST=$1 ALACGPU_QUAD_MIN_FIRST=$2 timeout 200 python bench.py --flags $f --steps 10 --warmup 3 --no-cpu --e2e-steps 3 --batch-tracks 0 > gpurun_out/q.json 2> gpurun_out/q.err || tail -3 gpurun_out/q.err
    python -c "
import json; d=json.load(open('gpurun_out/q.json')); s=d['stage_ms']; print('quads last>=$1 first>=$2 flags $f: value', round(d['value']), 'entropy %.3f lpc %.3f kernels %.3f e2e'%(s['entropy_ms'],s['lpc_ms'],s['kernels_ms']), round(d['e2e']['value']))"
  done
done
