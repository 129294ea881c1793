timeout 900 python -m pytest tests/test_gpu_step_parity.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | grep -v "^$" | tail -2
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2fM step %.4f' % (d['value']/1e6, d['ms_per_step']), {k: round(v, 4) for k, v in d['kernel_ms'].items()})"
MD_DEV_GRAPH=0 MD_LIB=build/libmdstep_clk.so python scripts/phase_clk.py cfg2 2>&1 | grep "k_post phase 2"
