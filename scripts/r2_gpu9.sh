set -x
timeout 900 python -m pytest tests/test_gpu_step_parity.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-other-configs 2>gpurun_out/err9.txt | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6), d['kernel_ms'])"
MD_LIB=build/libmdstep_clk.so python scripts/phase_clk.py cfg2 2>&1 | tail -22
FLUSH_READ=1 MD_LIB=build/libmdstep_clk.so python scripts/phase_clk.py cfg2 2>&1 | tail -22
