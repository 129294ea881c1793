# A/B of two builds of libmdstep.so on the bench workload: scripts/ab.sh libA.so libB.so
for l in $1 $2 $1 $2; do MD_LIB=$l timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$l value %.2fM' % (d['value']/1e6), {k: round(v, 4) for k, v in d['kernel_ms'].items()})"; done
