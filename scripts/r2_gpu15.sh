for mb in 4 6; do for t in 4 8 16; do MD_SCAN_TEAM=$t MD_SCAN_MB=$mb timeout 600 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-other-configs 2>gpurun_out/err15.txt | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mb=$mb team=$t value %.2fM' % (d['value']/1e6), {k: round(v, 4) for k, v in d['kernel_ms'].items()})"; done; done
