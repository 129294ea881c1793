# tollgate env on the device: parity with the oracle + the reference trace, env surface, and a short bench (no regression)
timeout 500 python -m pytest tests/test_gpu_step_parity.py tests/test_gpu_env_api.py -x -q -m gpu -k "tollgate or bottleneck or multi_agent_step" > gpurun_out/r2_gpu38.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_gpu38.log
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-other-configs > gpurun_out/r2_bench38.json 2> gpurun_out/r2_bench38.err
echo "bench rc $?" >> gpurun_out/r2_gpu38.log
