timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for f in 1 0 1 0; do MD_FUSE_BANK=$f timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-other-configs 2>gpurun_out/err20.txt | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse=$f value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6), {k: round(v, 4) for k, v in d['kernel_ms'].items()})"; done
