timeout 900 python -m pytest tests/test_gpu_step_parity.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline 2>gpurun_out/err17.txt | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6), {k: round(v, 4) for k, v in d['kernel_ms'].items()})
for o in d.get('other_configs', []): print(o.get('config', {}).get('baseline_config'), '%.2fM' % (o.get('value')/1e6), {k: round(v, 4) for k, v in o.get('kernel_ms').items()})"
MD_LIB=build/libmdstep_clk.so python scripts/phase_clk.py cfg2 2>&1 | grep -A10 "^k_pre:"
