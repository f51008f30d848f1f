#!/bin/bash
# round 2, visit T: final state -- whole suite, smoke, full bench (both arms)
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -12
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 800 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err
python -c "
import json
d=json.load(open('$O/bench.json')); r=d['roofline']; c=d['configs']
print('H', round(d['value']), round(d['ms_per_step']*1e3,1), 'frac', round(r['frac'],3), 'traffic', r.get('traffic'), 'launches', d['gpu_launches'])
print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in c.items() if 'value' in v}); print({k:(round(v['value']), round(v['ms_per_block'],3)) for k,v in c.get('e2e_dropin',{}).items() if isinstance(v,dict)})
print('e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'pipelined', round(d['e2e']['pipelined']['value']), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'clocks', d['clocks'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_ref.json')); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
