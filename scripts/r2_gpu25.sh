for p in 0 1 0 1; do if [ $p = 1 ]; then export BENCH_NO_PROF=1; else unset BENCH_NO_PROF; fi; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-other-configs 2>gpurun_out/err25.txt | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('noprof=$p value %.2fM step %.4f ms' % (d['value']/1e6, d['ms_per_step']))"; tail -2 gpurun_out/err25.txt; done
