set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; tail -c 600 gpurun_out/bench_r1_b.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_b_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_b.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_b.log 2>&1
for k in k1_entropy k2_lpc k3_stereo_pack; do
ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_r1b_$k python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_b_$k.log 2>&1
done
ls -la gpurun_out
