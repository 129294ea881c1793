# usage: scripts/tune_lib.sh "<lib suffixes under tune/>" "<workloads>": bench each tuning build of libmdstep on each workload
mkdir -p gpurun_out
for w in $2; do for lib in $1; do
  L=$PWD/tune/libmdstep_$lib.so; [ "$lib" = "main" ] && L=$PWD/metadrive_ped_b200/libmdstep.so
  MD_LIB=$L python bench.py --workload $w --steps 40 --warmup 5 --burnin 80 --no-cpu-baseline 2>&1 | grep '^{' \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print('$w $lib', '%.2fM/s' % (d['value']/1e6), ' '.join('%s=%.3f' % (n, v) for n, v in k.items()))"
done; done | tee -a gpurun_out/tune.log
