# usage: scripts/tune.sh "<lib suffixes>" "<MD_EPB_PRE values>" "<MD_EPB_POST values>" "<MD_EPB_DYN values>"
mkdir -p gpurun_out
for lib in $1; do for a in $2; do for b in $3; do for c in $4; do
  MD_LIB=$PWD/tune/libmdstep_$lib.so MD_EPB_PRE=$a MD_EPB_POST=$b MD_EPB_DYN=$c python bench.py --steps 40 --warmup 5 --burnin 80 --no-cpu-baseline  \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernel_ms']; print('$lib pre$a post$b dyn$c', '%.1fM/s' % (d['value']/1e6), ' '.join('%s=%.3f' % (n, v) for n, v in k.items()))"
done; done; done; done | tee -a gpurun_out/tune.log
