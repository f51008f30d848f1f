#!/bin/bash
# round 2, visit X: headline tests with the one-grid variant, default bench sanity after the last default change
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -6
timeout 200 python bench.py --no-cpu --no-e2e --no-kernels --only-configs C3,H_2048 --steps 200 > $O/x_default.json 2>/dev/null
python -c "
import json
d=json.load(open('$O/x_default.json')); c=d['configs']
print('H', round(d['value']), round(d['ms_per_step']*1e3,1), d['roofline']['plan'], {k:(round(v['value']), round(v['ms_per_block']*1e3,1)) for k,v in c.items() if 'value' in v})"
