#!/bin/bash
# round 2, visit L: bulk-copy cluster exchange staged in the buffers' own second halves (69.6 KB per CTA again) + L1-bypassing loads of the once-read rows
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
print('$2'.ljust(28), round(d['value']), round(d['ms_per_step']*1e3,1), 'us frac', round(r['frac'],3), 'B/s', round(r['algorithmic_bytes_per_sample'],1), {n[6:]:(round(v['alone_us'],1), round(v.get('alone_frac',0),2)) for n,v in k.items()})"; }
B="timeout 120 python bench.py --no-cpu --no-configs --no-e2e --steps 300"
run() { name=$1; shift; env "$@" $B $EXTRA > $O/l_$name.json 2>/dev/null; show $O/l_$name.json $name; }
EXTRA=""
run bulk_f0 DSP_B200_FIR_T2=0
run fenced_f0 DSP_B200_LIB=$OLD DSP_B200_FIR_T2=0
EXTRA="--no-kernels"
run bulk_f8u DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run fenced_f8u DSP_B200_LIB=$OLD DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run bulk_f0_nohot DSP_B200_FIR_T2=0 DSP_B200_FIR_HOT=0
EXTRA="--block 2048"
run b2048_bulk_f8u DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run b2048_fenced_f8u DSP_B200_LIB=$OLD DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
EXTRA="--no-kernels --block 2048"
run b2048_bulk_f8 DSP_B200_FIR_T2=8
run b2048_bulk_f0 DSP_B200_FIR_T2=0
EXTRA="--no-kernels --channels 64"
run c64_bulk_f0 DSP_B200_FIR_T2=0
run c64_fenced_f0 DSP_B200_LIB=$OLD DSP_B200_FIR_T2=0
