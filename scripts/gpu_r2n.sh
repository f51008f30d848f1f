#!/bin/bash
# round 2, visit N: far tier one near-period ahead (slack); high-priority stream and staggering on the config-5 chain
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -12
show() { python -c "
import json,sys
try: d=json.load(open('$1'))
except Exception as e: print('$2', 'no json', e); sys.exit(0)
r=d['roofline']
c=d.get('configs') or {}
print('$2'.ljust(28), round(d['value']), round(d['ms_per_step']*1e3,1), 'us frac', round(r['frac'],3), 'B/s', round(r['algorithmic_bytes_per_sample'],1), {k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in c.items() if 'value' in v})"; }
B="timeout 200 python bench.py --no-cpu --no-e2e --no-kernels --steps 200"
run() { name=$1; shift; env "$@" $B $EXTRA > $O/n_$name.json 2>/dev/null; show $O/n_$name.json $name; }
EXTRA="--only-configs C3,H_2048,C5_share"
run default X=1
run nohot DSP_B200_FIR_HOT=0
run nohot_unstag DSP_B200_FIR_HOT=0 DSP_B200_FIR_STAGGER=0
EXTRA="--only-configs H_2048,C5_share"
run hot_unstag DSP_B200_FIR_STAGGER=0
run nohot_f12 DSP_B200_FIR_HOT=0 DSP_B200_FIR_T2=12
EXTRA="--only-configs H_2048 --block 2048"
run b2048_default X=1
run b2048_nohot DSP_B200_FIR_HOT=0
