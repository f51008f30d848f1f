#!/bin/bash
# round 2, visit Z: MAC stream priority, staggered T = 6 / 8; then the final state once more (suite, smoke, bench)
mkdir -p gpurun_out
O=gpurun_out
one() { name=$1; shift; env "$@" timeout 100 python bench.py --no-cpu --no-e2e --no-kernels --no-configs --steps 300 > $O/z_$name.json 2>/dev/null; python -c "
import json
d=json.load(open('$O/z_$name.json')); print('$name'.ljust(16), round(d['value']), round(d['ms_per_step']*1e3,1), round(d['roofline']['algorithmic_bytes_per_sample'],1))"; }
one default X=1
one macprio_m2 DSP_B200_FIR_MAC_PRIO=-2
one macprio_m4 DSP_B200_FIR_MAC_PRIO=-4
one t6_stag DSP_B200_FIR_T=6
one t8_stag DSP_B200_FIR_T=8
one default_b X=1
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -6
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 800 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err
python -c "
import json
d=json.load(open('$O/bench.json')); r=d['roofline']; c=d['configs']
print('H', round(d['value']), round(d['ms_per_step']*1e3,1), 'frac', round(r['frac'],3), 'traffic', r.get('traffic'), 'launches', d['gpu_launches'])
print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in c.items() if 'value' in v}); print({k:(round(v['value']), round(v['ms_per_block'],3)) for k,v in c.get('e2e_dropin',{}).items() if isinstance(v,dict)})
print('e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'pipelined', round(d['e2e']['pipelined']['value']), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
