# usage: scripts/tune_env.sh VAR1 "<values>" VAR2 "<values>" [extra bench args]: sweeps two env-var launch parameters of libmdstep
mkdir -p gpurun_out
V1=$1; L1=$2; V2=$3; L2=$4; shift 4
for a in $L1; do for b in $L2; do
  env $V1=$a $V2=$b python bench.py --steps 40 --warmup 5 --burnin 80 --no-cpu-baseline "$@" 2>&1 | grep '^{' \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print('$V1=$a $V2=$b', '%.2fM/s' % (d['value']/1e6), ' '.join('%s=%.3f' % (n, v) for n, v in k.items()))"
done; done | tee -a gpurun_out/tune.log
