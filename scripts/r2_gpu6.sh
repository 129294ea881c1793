set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r2_b6.json 2> gpurun_out/r2_b6.err; tail -3 gpurun_out/r2_b6.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_b6.json').read().strip().splitlines()[-1])
print('value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6), d['kernel_ms'])
for o in d.get('other_configs', []): print(o.get('config', {}).get('baseline_config'), o.get('value'), o.get('kernel_ms'))
PY
MD_LIB=build/libmdstep_clk.so python scripts/phase_clk.py cfg2 2>&1 | tail -22
