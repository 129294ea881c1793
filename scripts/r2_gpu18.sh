timeout 900 python -m pytest tests/test_gpu_step_parity.py -x -q -m gpu 2>&1 | tail -2
for l in metadrive_ped_b200/libmdstep.so build/libmdstep_nopf.so metadrive_ped_b200/libmdstep.so build/libmdstep_nopf.so; do MD_LIB=$l timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-other-configs 2>gpurun_out/err18.txt | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$l value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6), {k: round(v, 4) for k, v in d['kernel_ms'].items()})"; done
