for g in 0 1 0 1; do MD_EXP_NO_CONTACTS=$g timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nocontacts=$g value %.2fM step %.4f' % (d['value']/1e6, d['ms_per_step']), {k: round(v, 4) for k, v in d['kernel_ms'].items()})"; done
