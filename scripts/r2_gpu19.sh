timeout 900 python -m pytest tests/test_gpu_step_parity.py -x -q -m gpu -k "degenerate" 2>&1 | tail -25
