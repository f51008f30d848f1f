#!/bin/bash
# round 2, visit H: software-pipelined batch kernel (variant library) against the default, per-kernel times of the tier forms
mkdir -p gpurun_out
O=gpurun_out
PF=$PWD/dsp_b200/variants/libdspb200_pf.so
DSP_B200_LIB=$PF timeout 400 python -m pytest tests/test_gpu_headline.py -m gpu -q --no-header -p no:cacheprovider -k "composition or 2048 or ragged" > $O/pytest_pf.log 2>&1
echo "pytest exit $?" >> $O/pytest_pf.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_pf.log | tail -12
show() { python -c "
import json,sys
try: d=json.load(open('$1'))
except Exception as e: print('$2', 'no json', e); sys.exit(0)
r=d['roofline']; k=r.get('kernels') or {}
print('$2'.ljust(28), round(d['value']), round(d['ms_per_step']*1e3,1), 'us frac', round(r['frac'],3), 'B/s', round(r['algorithmic_bytes_per_sample'],1), {n[6:]:(round(v['alone_us'],1), round(v.get('alone_frac',0),2)) for n,v in k.items()})"; }
B="timeout 120 python bench.py --no-cpu --no-configs --no-e2e --steps 300"
run() { name=$1; shift; env "$@" $B $EXTRA > $O/h_$name.json 2>/dev/null; show $O/h_$name.json $name; }
EXTRA=""
run f0 DSP_B200_FIR_T2=0
run f8u DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run pf_f0 DSP_B200_LIB=$PF DSP_B200_FIR_T2=0
run pf_f8 DSP_B200_LIB=$PF DSP_B200_FIR_T2=8
run pf_f8u DSP_B200_LIB=$PF DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run pf_f12u DSP_B200_LIB=$PF DSP_B200_FIR_T2=12 DSP_B200_FIR_STAGGER=0
run pf_t8_f0 DSP_B200_LIB=$PF DSP_B200_FIR_T2=0 DSP_B200_FIR_T=8
EXTRA="--block 2048"
run b2048_f8 DSP_B200_FIR_T2=8
run b2048_f8u DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run b2048_pf_f8 DSP_B200_LIB=$PF DSP_B200_FIR_T2=8
run b2048_pf_f8u DSP_B200_LIB=$PF DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run b2048_pf_f12 DSP_B200_LIB=$PF DSP_B200_FIR_T2=12
run b2048_pf_f16 DSP_B200_LIB=$PF DSP_B200_FIR_T2=16
