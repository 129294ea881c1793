set -x
nvidia-smi -L
python -c "import torch; print(torch.cuda.is_available(), torch.cuda.get_device_name(0))"
python -m pytest tests/test_gpu_step_parity.py -x -q -m gpu 2>&1 | tail -40
