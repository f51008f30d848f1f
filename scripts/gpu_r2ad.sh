#!/bin/bash
# round 2, visit AD: the full bench line on the final state (with the device block cache)
mkdir -p gpurun_out
O=gpurun_out
timeout 800 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err
python -c "
import json
d=json.load(open('$O/bench.json')); r=d['roofline']; c=d['configs']
print('H', round(d['value']), round(d['ms_per_step']*1e3,1), 'frac', round(r['frac'],3), 'traffic', r.get('traffic'), 'launches', d['gpu_launches'])
print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in c.items() if 'value' in v}, c['H_2048'].get('runs_ms_per_block')); print({k:(round(v['value']), round(v['ms_per_block'],3)) for k,v in c.get('e2e_dropin',{}).items() if isinstance(v,dict)})
print('e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'pipelined', round(d['e2e']['pipelined']['value']), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks'])"
