set -x
timeout 900 python -m pytest tests/test_gpu_host_path.py tests/test_gpu_env_api.py -x -q -m gpu 2>&1 | tail -5
for g in 1 2 3 4 6 8; do timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --host-groups $g > gpurun_out/r2_bench_cfg2_g$g.json 2>gpurun_out/err.txt || tail -5 gpurun_out/err.txt; done
MD_HOST_GRAPH=0 timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --host-groups 4 > gpurun_out/r2_bench_cfg2_g4_nograph.json 2>/dev/null
timeout 600 python bench.py --steps 50 --warmup 10 --workload cfg3 --no-cpu-baseline > gpurun_out/r2_bench_cfg3.json 2> gpurun_out/r2_bench_cfg3.err; tail -3 gpurun_out/r2_bench_cfg3.err
cat gpurun_out/r2_bench_cfg2_g*.json gpurun_out/r2_bench_cfg3.json | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d['config'].get('baseline_config'), 'value %.2fM' % (d['value']/1e6), 'e2e', {k: (round(v/1e6,2) if isinstance(v, float) and v > 1e4 else v) for k, v in d['e2e'].items() if k != 'api'})
"
