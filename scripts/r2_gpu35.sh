for e in 8 16 12 6; do MD_EPB_DYN=$e timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('epb_dyn=$e value %.2fM step %.4f' % (d['value']/1e6, d['ms_per_step']), {k: round(v, 4) for k, v in d['kernel_ms'].items()})"; done
