#!/bin/bash
# round 2, visit AA: far tier in two turns (half the channels every fourth block) against whole launches, 2048-frame blocks, repeated
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -6
one() { name=$1; shift; env "$@" timeout 100 python bench.py --no-cpu --no-e2e --no-kernels --no-configs --block 2048 --steps 300 > $O/aa_$name.json 2>/dev/null; python -c "
import json
d=json.load(open('$O/aa_$name.json')); print('$name'.ljust(16), round(d['value']), round(d['ms_per_step']*1e3,1), round(d['roofline']['algorithmic_bytes_per_sample'],1))"; }
for r in 1 2 3 4; do one k2_$r X=1; one k1_$r DSP_B200_FIR_FAR_CLASSES=1; done
one k4_e8 DSP_B200_FIR_FAR_CLASSES=1 DSP_B200_FIR_T2=16
timeout 200 python bench.py --no-cpu --no-e2e --no-kernels --only-configs C3,H_2048 --steps 100 > $O/aa_cfg.json 2>/dev/null
python -c "
import json
d=json.load(open('$O/aa_cfg.json')); c=d['configs']
print({k:(round(v['value']), round(v['ms_per_block']*1e3,1), v.get('runs_ms_per_block')) for k,v in c.items() if 'value' in v})"
