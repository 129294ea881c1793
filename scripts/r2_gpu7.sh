set -x
MD_DEBUG_OCC=1 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-other-configs > gpurun_out/r2_b7.json 2> gpurun_out/r2_b7.err; grep mdstep gpurun_out/r2_b7.err | sort | uniq -c; tail -3 gpurun_out/r2_b7.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_b7.json').read().strip().splitlines()[-1])
print('value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6), d['kernel_ms'])
PY
MD_LIB=build/libmdstep_clk.so python scripts/phase_clk.py cfg2 2>&1 | tail -22
for t in 64 128 160; do MD_DYN_THREADS=$t timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads $t', 'value %.2fM' % (d['value']/1e6), d['kernel_ms'])"; done
