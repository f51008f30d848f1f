#!/bin/bash
# round 2, visit G: packed-bin special case only in the warp that holds it; fused kernel capped at one CTA per SM; thread counts
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -12
show() { python -c "
import json,sys
try: d=json.load(open('$1'))
except Exception as e: print('$2', 'no json', e); sys.exit(0)
r=d['roofline']; k=r.get('kernels') or {}
print('$2'.ljust(28), round(d['value']), round(d['ms_per_step']*1e3,1), 'us frac', round(r['frac'],3), 'B/s', round(r['algorithmic_bytes_per_sample'],1), {n[6:]:(round(v['alone_us'],1), round(v.get('alone_frac',0),2)) for n,v in k.items()})"; }
B="timeout 120 python bench.py --no-cpu --no-configs --no-e2e --steps 300"
run() { name=$1; shift; env "$@" $B $EXTRA > $O/g_$name.json 2>/dev/null; show $O/g_$name.json $name; }
EXTRA=""
run t4_f12 X=1
run t4_f8 DSP_B200_FIR_T2=8
EXTRA="--no-kernels"
run t4_f0 DSP_B200_FIR_T2=0
run t4_f0_unstag DSP_B200_FIR_T2=0 DSP_B200_FIR_STAGGER=0
run t4_f8_unstag DSP_B200_FIR_T2=8 DSP_B200_FIR_STAGGER=0
run l0one_f0 DSP_B200_FIR_T2=0 DSP_B200_FIR_L0_SMEM_KB=120
run l0one_f0_m128 DSP_B200_FIR_T2=0 DSP_B200_FIR_L0_SMEM_KB=120 DSP_B200_FIR_MAC_THREADS=128 DSP_B200_FIR_BATCH_THREADS=128
run l0one_f8 DSP_B200_FIR_T2=8 DSP_B200_FIR_L0_SMEM_KB=120
run l0one_f8_m128 DSP_B200_FIR_T2=8 DSP_B200_FIR_L0_SMEM_KB=120 DSP_B200_FIR_MAC_THREADS=128 DSP_B200_FIR_BATCH_THREADS=128
run l0one_f12_m128 DSP_B200_FIR_L0_SMEM_KB=120 DSP_B200_FIR_MAC_THREADS=128 DSP_B200_FIR_BATCH_THREADS=128
run f0_m128 DSP_B200_FIR_T2=0 DSP_B200_FIR_MAC_THREADS=128 DSP_B200_FIR_BATCH_THREADS=128
EXTRA="--no-kernels --block 2048"
run b2048_f8 DSP_B200_FIR_T2=8
run b2048_f12 X=1
run b2048_f16 DSP_B200_FIR_T2=16
run b2048_2lv DSP_B200_FIR_SINGLE_MIN=4096
run b2048_f8_m128 DSP_B200_FIR_T2=8 DSP_B200_FIR_MAC_THREADS=128 DSP_B200_FIR_BATCH_THREADS=128
