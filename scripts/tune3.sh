# k_pre sweep: lib suffixes x MD_EPB_PRE x MD_PRE_WORKERS
mkdir -p gpurun_out
for lib in $1; do for b in $2; do for w in $3; do
  MD_LIB=$PWD/tune/libmdstep_$lib.so MD_EPB_PRE=$b MD_PRE_WORKERS=$w python bench.py --steps 40 --warmup 5 --burnin 80 --no-cpu-baseline 2>&1 | grep '^{' \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print('$lib pre$b workers$w', '%.1fM/s' % (d['value']/1e6), ' '.join('%s=%.3f' % (n, v) for n, v in k.items()))"
done; done; done | tee -a gpurun_out/tune.log
