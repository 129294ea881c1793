set -x
N=${1:-8}
nvidia-smi topo -m 2>/dev/null | head -14
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -3 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('N=%d value %.2fM e2e %.2fM' % (d['n_gpus'], d['value']/1e6, d['e2e']['value']/1e6), d['kernel_ms'])
for o in d.get('other_configs', []): print(o.get('config', {}).get('baseline_config'), '%.2fM' % (o.get('value')/1e6), 'e2e %.2fM' % (o.get('e2e', {}).get('value', 0)/1e6))
PY
