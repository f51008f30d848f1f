#!/bin/bash
# round 2, visit J: fused kernel's cluster exchange by bulk copies (no fences) and on a highest-priority stream, against the fenced form / the caller's stream
mkdir -p gpurun_out
O=gpurun_out
OLD=$PWD/dsp_b200/variants/libdspb200_l0sync.so
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -12
show() { python -c "
import json,sys
try: d=json.load(open('$1'))
except Exception as e: print('$2', 'no json', e); sys.exit(0)
r=d['roofline']; k=r.get('kernels') or {}
print('$2'.ljust(28), round(d['value']), round(d['ms_per_step']*1e3,1), 'us frac', round(r['frac'],3), 'B/s', round(r['algorithmic_bytes_per_sample'],1), {n[6:]:(round(v['alone_us'],1), round(v.get('alone_frac',0),2)) for n,v in k.items()})
c=d.get('configs') or {}
if c: print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in c.items() if 'value' in v})"; }
B="timeout 120 python bench.py --no-cpu --no-configs --no-e2e --steps 300"
run() { name=$1; shift; env "$@" $B $EXTRA > $O/j_$name.json 2>/dev/null; show $O/j_$name.json $name; }
EXTRA=""
run bulk_hot_f0 DSP_B200_FIR_T2=0
EXTRA="--no-kernels"
run bulk_nohot_f0 DSP_B200_FIR_T2=0 DSP_B200_FIR_HOT=0
run fenced_hot_f0 DSP_B200_LIB=$OLD DSP_B200_FIR_T2=0
run fenced_nohot_f0 DSP_B200_LIB=$OLD DSP_B200_FIR_T2=0 DSP_B200_FIR_HOT=0
run bulk_hot_f0_unstag DSP_B200_FIR_T2=0 DSP_B200_FIR_STAGGER=0
run bulk_hot_f8u DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run bulk_hot_f8 DSP_B200_FIR_T2=8
run bulk_hot_f12u DSP_B200_FIR_T2=12 DSP_B200_FIR_STAGGER=0
run bulk_hot_f12 X=1
EXTRA="--no-kernels --block 2048"
run b2048_bulk_hot_f8u DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run b2048_bulk_hot_f8 DSP_B200_FIR_T2=8
run b2048_bulk_hot_f16u DSP_B200_FIR_T2=16 DSP_B200_FIR_STAGGER=0
run b2048_bulk_hot_f12 X=1
run b2048_fenced_hot_f8u DSP_B200_LIB=$OLD DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run b2048_bulk_nohot_f8u DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0 DSP_B200_FIR_HOT=0
run b2048_2lv DSP_B200_FIR_SINGLE_MIN=4096
EXTRA="--no-kernels --channels 64"
run c64_bulk_hot_f0 DSP_B200_FIR_T2=0
run c64_bulk_hot_f8u DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
