#!/bin/bash
# usage: profiles/sweep.sh <tag> <env assignments...> ; runs bench cfg2+cfg3 quickly and prints the key numbers
tag=$1; shift
for w in ${WL:-cfg2 cfg3}; do
  env "$@" timeout 300 python bench.py --workload $w --steps 2000 --warmup 20 --cpu-seconds 0.3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if not l.startswith('{'): continue
    j=json.loads(l); r=j['roofline']
    print('$tag', '$w', 'Gbp/s=%.1f'%(j['value']/1e9), 'step_us=%.2f'%(j['ms_per_step']*1e3), 'exec_us=%.2f'%(r['launch_ms']*1e3), 'exec1s_us=%.2f'%(r.get('launch_ms_one_stream',0)*1e3), 'plan_us=%.2f'%(r['plan_kernel_ms']*1e3), 'iso_us=%.2f'%(r['launch_ms_isolated_after_l2_flush']*1e3), 'frac=%.3f'%r['frac'], 'step_frac=%.3f'%r['whole_step_frac'])
"
done
