set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -2 gpurun_out/r02_bench_1gpu.err
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -2 gpurun_out/r02_bench_reference.err
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-other-configs --host-groups 2 > gpurun_out/r02_bench_g2.json 2>/dev/null
bash scripts/gpu_profile.sh > gpurun_out/profile.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
