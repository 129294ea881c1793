timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2_b13.json 2>gpurun_out/err13.txt; tail -2 gpurun_out/err13.txt
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_b13.json').read().strip().splitlines()[-1])
print('value %.2fM e2e %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6), {k: round(v, 4) for k, v in d['kernel_ms'].items()})
for o in d.get('other_configs', []): print(o.get('config', {}).get('baseline_config'), '%.2fM' % (o.get('value')/1e6), 'e2e', o.get('e2e', {}).get('value'), {k: round(v, 4) for k, v in o.get('kernel_ms').items()})
PY
python - <<'PY'
import torch, time
x = torch.empty(64*1024*1024, dtype=torch.uint8, device='cuda'); h = torch.empty(64*1024*1024, dtype=torch.uint8).pin_memory()
for name, (a, b) in {'d2h': (h, x), 'h2d': (x, h)}.items():
    for _ in range(3): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); t = time.time()
    for _ in range(20): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.time() - t
    print(name, '%.1f GB/s' % (20 * 64 * 1.048576e-3 / dt * 1.0))
PY
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
MD_LIB=build/libmdstep_clk.so python scripts/phase_clk.py cfg2 2>&1 | head -8
