set -x
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_full.log 2>&1
grep -n "Fatal\|Segmentation\|passed\|failed\|^tests/" gpurun_out/r2_pytest_full.log | head -20
grep -n "File \"/root/repo\|File \"/tmp\|line [0-9]* in test_" gpurun_out/r2_pytest_full.log | head -20
