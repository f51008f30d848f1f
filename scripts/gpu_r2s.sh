#!/bin/bash
# round 2, visit S: K3 input rows requested deeper ahead (tiles 5/6/7) against tile 4; host-mode calls without the priority hop
mkdir -p gpurun_out
O=gpurun_out
for t in 6 7; do DSP_B200_RS_TILE=$t timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q --no-header -p no:cacheprovider -k "resample or golden" > $O/pytest_rs$t.log 2>&1; echo "tile $t pytest exit $?"; grep -E 'passed|failed' $O/pytest_rs$t.log | tail -1; done
for t in 4 5 6 7; do
DSP_B200_RS_TILE=$t timeout 150 python bench.py --no-cpu --no-e2e --no-kernels --only-configs C4 --steps 50 > $O/s_tile$t.json 2>/dev/null
python -c "
import json
d=json.load(open('$O/s_tile$t.json')); c=d['configs']
print('tile $t', {k:(round(v['value']), round(v['ms_per_block']*1e3,1)) for k,v in c.items() if 'value' in v}, c['C4'].get('fp64',{}).get('frac_of_fp64_peak'))"
done
timeout 200 python bench.py --no-cpu --no-kernels --no-configs --steps 200 > $O/s_e2e.json 2>/dev/null
python -c "
import json
d=json.load(open('$O/s_e2e.json')); print('H', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'pipelined', round(d['e2e']['pipelined']['value']))"
