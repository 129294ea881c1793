timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^$" | tail -4
for g in 1 0 1 0; do MD_DEV_GRAPH=$g timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graph=$g value %.2fM step %.4f' % (d['value']/1e6, d['ms_per_step']), 'e2e %.2fM' % (d['e2e']['value']/1e6))"; done
