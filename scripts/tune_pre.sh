# usage: scripts/tune_pre.sh "<lib suffixes>" "<MD_PRE_WORKERS>" "<MD_PRE_TEAM>" "<workloads>"
mkdir -p gpurun_out
for w in $4; do for lib in $1; do for t in $2; do for tm in $3; do
  L=$PWD/tune/libmdstep_$lib.so; [ "$lib" = "main" ] && L=$PWD/metadrive_ped_b200/libmdstep.so
  MD_LIB=$L MD_PRE_WORKERS=$t MD_PRE_TEAM=$tm python bench.py --workload $w --steps 40 --warmup 5 --burnin 80 --no-cpu-baseline 2>&1 | grep '^{' \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print('$w $lib workers=$t team=$tm', '%.2fM/s' % (d['value']/1e6), ' '.join('%s=%.3f' % (n, v) for n, v in k.items()))"
done; done; done; done | tee -a gpurun_out/tune.log
