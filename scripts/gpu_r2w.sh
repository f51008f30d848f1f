#!/bin/bash
# round 2, visit W: per-block MAC + staggered batch class as one grid (k_fir_tail) against the two launches
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -12
show() { python -c "
import json,sys
try: d=json.load(open('$1'))
except Exception as e: print('$2', 'no json', e); sys.exit(0)
r=d['roofline']; k=r.get('kernels') or {}
c=d.get('configs') or {}
print('$2'.ljust(20), round(d['value']), round(d['ms_per_step']*1e3,1), 'us frac', round(r['frac'],3), 'launches', d.get('gpu_launches'), {n[6:]:(round(v['alone_us'],1), round(v.get('alone_frac',0),2)) for n,v in k.items()}, {k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1)) for k,v in c.items() if 'value' in v})"; }
B="timeout 200 python bench.py --no-cpu --no-e2e --steps 300"
run() { name=$1; shift; env "$@" $B $EXTRA > $O/w_$name.json 2>/dev/null; show $O/w_$name.json $name; }
EXTRA="--only-configs C3"
run merged X=1
run separate DSP_B200_FIR_MERGE=0
EXTRA="--no-kernels --no-configs"
run merged_b X=1
run separate_b DSP_B200_FIR_MERGE=0
run merged_nohot DSP_B200_FIR_HOT=0
EXTRA="--no-kernels --no-configs --block 8192 --taps 262144"
run b8192_merged X=1
run b8192_separate DSP_B200_FIR_MERGE=0
