#!/bin/bash
# round 2, visit K: is the bulk-copy form slower because of its 104 KB of shared memory per CTA? (fenced form with the same footprint)
mkdir -p gpurun_out
O=gpurun_out
OLD=$PWD/dsp_b200/variants/libdspb200_l0sync.so
show() { python -c "
import json,sys
try: d=json.load(open('$1'))
except Exception as e: print('$2', 'no json', e); sys.exit(0)
r=d['roofline']; k=r.get('kernels') or {}
print('$2'.ljust(28), round(d['value']), round(d['ms_per_step']*1e3,1), 'us frac', round(r['frac'],3), 'B/s', round(r['algorithmic_bytes_per_sample'],1), {n[6:]:(round(v['alone_us'],1), round(v.get('alone_frac',0),2)) for n,v in k.items()})"; }
B="timeout 120 python bench.py --no-cpu --no-configs --no-e2e --steps 300"
run() { name=$1; shift; env "$@" $B $EXTRA > $O/k_$name.json 2>/dev/null; show $O/k_$name.json $name; }
EXTRA=""
run fenced_103k DSP_B200_LIB=$OLD DSP_B200_FIR_T2=0 DSP_B200_FIR_L0_SMEM_KB=103
run fenced_86k DSP_B200_LIB=$OLD DSP_B200_FIR_T2=0 DSP_B200_FIR_L0_SMEM_KB=86
run fenced DSP_B200_LIB=$OLD DSP_B200_FIR_T2=0
