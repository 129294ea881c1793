set -x
nvidia-smi -L
nproc
lscpu | grep -i "numa\|model name\|socket" | head
nvidia-smi topo -m 2>/dev/null | head -20
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err; tail -3 gpurun_out/r2_bench_cfg2.err
timeout 600 python bench.py --steps 50 --warmup 10 --workload cfg3 --no-cpu-baseline > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; tail -3 gpurun_out/r2_bench_cfg3.err
for g in 1 2 8; do timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --host-groups $g > gpurun_out/r2_bench_cfg2_g$g.json 2>/dev/null; done
cat gpurun_out/r2_bench_cfg2*.json gpurun_out/r2_bench_cfg3.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['config'].get('baseline_config'), 'value %.2fM' % (d['value']/1e6), 'e2e', {k: (round(v/1e6,2) if isinstance(v, float) and v > 1e4 else v) for k, v in d['e2e'].items() if k != 'api'}, d['kernel_ms'])
"
