# usage: scripts/tune_team.sh "<MD_POST_TEAM values>" "<workloads>"
mkdir -p gpurun_out
for w in $2; do for t in $1; do
  MD_POST_TEAM=$t python bench.py --workload $w --steps 40 --warmup 5 --burnin 80 --no-cpu-baseline 2>&1 | grep '^{' \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print('$w team=$t', '%.2fM/s' % (d['value']/1e6), ' '.join('%s=%.3f' % (n, v) for n, v in k.items()))"
done; done | tee -a gpurun_out/tune.log
