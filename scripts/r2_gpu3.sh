set -x
timeout 600 python scripts/ma_driver_sweep.py 2>&1 | tail -12
