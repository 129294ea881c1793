for g in 1 2 3 4; do timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-other-configs --host-groups $g 2>/dev/null | python -c "
import json,sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d['e2e']; print('groups $g value %.2fM e2e %.2fM sync %.2fM' % (d['value']/1e6, e['value']/1e6, e['sync_call_value']/1e6))"; done
