# usage: bash tools/gpu_prof.sh <tag> <kernel-regex>...   (after the plain bench exited 0)
tag=$1; shift
python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/plain_$tag.json 2> gpurun_out/plain_$tag.err || { tail -5 gpurun_out/plain_$tag.err; exit 1; }
for k in "$@"; do
ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_${tag}_$k python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_${tag}_$k.log 2>&1
done
ls -la gpurun_out | tail -5
