timeout 900 python -m pytest tests/test_gpu_step_parity.py -x -q -m gpu 2>&1 | tail -2
for t in 8 16 32 4; do for c in 592 1184; do MD_SCAN_TEAM=$t MD_SCAN_CTAS=$c timeout 600 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-other-configs 2>gpurun_out/err12.txt | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('team=$t ctas=$c value %.2fM' % (d['value']/1e6), {k: round(v, 4) for k, v in d['kernel_ms'].items()})"; done; done
