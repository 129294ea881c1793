timeout 900 python -m pytest tests/test_gpu_step_parity.py -x -q -m gpu 2>&1 | tail -2
for mb in 4 6; do for t in 4 8; do MD_SCAN_TEAM=$t MD_SCAN_MB=$mb timeout 600 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-other-configs 2>gpurun_out/err16.txt | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mb=$mb team=$t value %.2fM' % (d['value']/1e6), {k: round(v, 4) for k, v in d['kernel_ms'].items()})"; done; done
MD_LIB=build/libmdstep_clk.so python scripts/phase_clk.py cfg2 2>&1 | grep -A6 k_scan
