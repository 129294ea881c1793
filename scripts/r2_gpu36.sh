run() { timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-other-configs 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 value %.2fM step %.4f' % (d['value']/1e6, d['ms_per_step']), {k: round(v, 4) for k, v in d['kernel_ms'].items() if k in ('k_pre', 'k_post')})"; }
run base
MD_EPB_POST=8 run post_epb8
MD_EPB_POST=32 MD_POST_WORKERS=256 run post_epb32_w256
MD_EPB_POST=16 MD_POST_WORKERS=64 run post_epb16_w64
MD_EPB_POST=8 MD_POST_WORKERS=64 run post_epb8_w64
MD_EPB_PRE=8 MD_PRE_WORKERS=128 run pre_epb8_w128
MD_EPB_PRE=32 MD_PRE_WORKERS=512 run pre_epb32_w512
MD_EPB_PRE=16 MD_PRE_WORKERS=192 run pre_epb16_w192
