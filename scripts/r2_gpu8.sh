for r in 80 72 64; do MD_DEBUG_OCC=1 MD_LIB=build/libmdstep_dyn$r.so timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs 2>gpurun_out/err8.txt | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('regs $r', 'value %.2fM' % (d['value']/1e6), d['kernel_ms'])"; grep "k_dyn: 1024" gpurun_out/err8.txt; done
