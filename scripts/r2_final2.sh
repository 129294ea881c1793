# closing GPU session after the tollgate env and the top-down observation: the whole GPU suite, the bench line, smoke
set -x
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_full2.log 2>&1; tail -3 gpurun_out/r2_pytest_full2.log
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu_b.json 2> gpurun_out/r02_bench_1gpu_b.err; tail -2 gpurun_out/r02_bench_1gpu_b.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
