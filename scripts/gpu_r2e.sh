#!/bin/bash
# round 2, visit E: suite, smoke, bench twice (variance of the configs), batch-depth sweep on the default path
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error' $O/pytest_gpu.log | tail -20
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -3 $O/smoke.log
show() { python -c "
import json; d=json.load(open('$1')); print('$2 H', round(d['value']), round(d['ms_per_step']*1e3,1), round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None)
c=d.get('configs') or {}
print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in c.items() if 'value' in v})
print({k:(round(v['value']), round(v['ms_per_block'],3)) for k,v in c.get('e2e_dropin',{}).items() if isinstance(v,dict)})
cp=d.get('cpu_baseline'); print('cpu', cp['value'] if cp else None)"; }
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err; show $O/bench.json first
timeout 400 python bench.py --no-cpu > $O/bench2.json 2> $O/bench2.err; show $O/bench2.json second
for t in 6 8; do DSP_B200_FIR_T=$t timeout 200 python bench.py --no-cpu --no-configs --no-e2e > $O/bench_t$t.json 2>/dev/null; show $O/bench_t$t.json T$t; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_ref.json')); print('reference arm', d['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['spread'])"
