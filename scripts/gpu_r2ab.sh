#!/bin/bash
# round 2, visit AB: headline tests on the final defaults; the 2048-frame config measured after the 4096-frame headline in the same process
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -6
cfg() { name=$1; shift; env "$@" timeout 150 python bench.py --no-cpu --no-e2e --no-kernels --only-configs H_2048 --steps 100 $EXTRA > $O/ab_$name.json 2>/dev/null; python -c "
import json
d=json.load(open('$O/ab_$name.json')); c=d['configs']['H_2048']
print('$name'.ljust(18), round(d['ms_per_step']*1e3,1), [round(x*1e3,1) for x in c.get('runs_ms_per_block',[])])"; }
EXTRA=""
cfg after4096 X=1
cfg after4096_nohot DSP_B200_FIR_HOT=0
cfg after4096_f0 DSP_B200_FIR_T2=0
cfg after4096_2lv DSP_B200_FIR_SINGLE_MIN=4096
EXTRA="--block 2048"
cfg after2048 X=1
EXTRA="--channels 64"
cfg after64ch X=1
