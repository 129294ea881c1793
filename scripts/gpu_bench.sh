set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
for w in cfg3 cfg4 cfg5; do python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -c 1500 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err; done
nproc
