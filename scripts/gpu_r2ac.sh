#!/bin/bash
# round 2, visit AC: device block cache -- whole suite, smoke, the 2048-frame config (three fresh chains) after the 4096-frame headline, cache on / off
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -6
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
cfg() { name=$1; shift; env "$@" timeout 150 python bench.py --no-cpu --no-e2e --no-kernels --only-configs C3,H_2048,C5_share --steps 100 > $O/ac_$name.json 2>/dev/null; python -c "
import json
d=json.load(open('$O/ac_$name.json')); c=d['configs']
print('$name'.ljust(12), round(d['ms_per_step']*1e3,1), [round(x*1e3,1) for x in c['H_2048'].get('runs_ms_per_block',[])], {k:round(v['ms_per_block']*1e3,1) for k,v in c.items() if 'value' in v})"; }
cfg pool_1 X=1
cfg pool_2 X=1
cfg nopool_1 DSP_B200_POOL_MB=0
cfg pool_3 X=1
