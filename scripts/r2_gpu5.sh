set -x
timeout 900 python -m pytest tests/test_gpu_step_parity.py tests/test_gpu_env_api.py -x -q -m gpu -k "bottleneck or multi_agent or ma_" 2>&1 | tail -15
