timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -3
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.2fM step %.4f e2e %.2fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6), {k: round(v, 4) for k, v in d['kernel_ms'].items()})
for o in d.get('other_configs', []): print('   ', o.get('config', {}).get('baseline_config'), '%.2fM' % (o.get('value')/1e6), 'e2e %.2fM' % (o.get('e2e', {}).get('value', 0)/1e6), {k: round(v, 4) for k, v in o['kernel_ms'].items()})"
