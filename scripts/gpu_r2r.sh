#!/bin/bash
# round 2, visit R: K3 with the tap tile double-buffered (cp.async) against the single-buffered form
mkdir -p gpurun_out
O=gpurun_out
DSP_B200_RS_TILE=5 timeout 400 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "resample or golden or chain or dropin" > $O/pytest_rs.log 2>&1
echo "pytest exit $?" >> $O/pytest_rs.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_rs.log | tail -8
for t in 4 5 4 5; do
DSP_B200_RS_TILE=$t timeout 150 python bench.py --no-cpu --no-e2e --no-kernels --only-configs C4,C5_share --steps 50 > $O/r_tile$t.json 2>/dev/null
python -c "
import json
d=json.load(open('$O/r_tile$t.json')); c=d['configs']
print('tile $t', {k:(round(v['value']), round(v['ms_per_block']*1e3,1)) for k,v in c.items() if 'value' in v}, c['C4'].get('fp64',{}).get('frac_of_fp64_peak'))"
done
