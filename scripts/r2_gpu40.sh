timeout 500 python -m pytest tests/test_topdown.py -x -q -m gpu > gpurun_out/r2_gpu40.log 2>&1
echo "pytest rc $?" >> gpurun_out/r2_gpu40.log
PYTHONPATH=. timeout 300 python scripts/topdown_time.py > gpurun_out/r2_topdown_time.json 2>> gpurun_out/r2_gpu40.log
echo "time rc $?" >> gpurun_out/r2_gpu40.log
