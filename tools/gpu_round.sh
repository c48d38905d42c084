# Round-end evidence run (one GPU): tests, smoke, bench (both arms), ncu launch list + full captures.
# usage: bash tools/gpu_round.sh <tag>
tag=${1:-r1}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || { tail -5 gpurun_out/bench_${tag}.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --batch-tracks 0 > gpurun_out/plain_${tag}.json 2> gpurun_out/plain_${tag}.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --batch-tracks 0 > gpurun_out/ncu_${tag}.log 2>&1
for k in k123_decode k12_entropy_lpc k3_stereo_pack; do
skip=2; [ $k = k123_decode ] && skip=0      # the fully fused launch runs once (zero-copy parity pass)
ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip --launch-count 1 -f -o gpurun_out/prof_${tag}_$k python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --batch-tracks 0 > gpurun_out/ncu_${tag}_$k.log 2>&1
done
ls -la gpurun_out | tail -12
