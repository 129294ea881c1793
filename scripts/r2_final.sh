# the round's closing GPU session: tests, bench (both arms), ncu captures, phase clocks, smoke -> gpurun_out/ (then profiles/)
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -2 gpurun_out/r02_bench_1gpu.err
timeout 900 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -2 gpurun_out/r02_bench_reference.err
bash scripts/gpu_profile.sh > gpurun_out/profile.log 2>&1
MD_DEV_GRAPH=0 MD_LIB=build/libmdstep_clk.so python scripts/phase_clk.py cfg2 > gpurun_out/r02_phase_clock.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
