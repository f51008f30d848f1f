#!/bin/bash
# round 2, visit F: two-tier staggered batched tail + single-level 2048-frame blocks: suite, smoke, bench, tier sweep
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -20
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
show() { python -c "
import json,sys
try: d=json.load(open('$1'))
except Exception as e: print('$2', 'no json', e); sys.exit(0)
r=d['roofline']; print('$2 H', round(d['value']), round(d['ms_per_step']*1e3,1), 'us  frac', round(r['frac'],3), 'B/sample', round(r['algorithmic_bytes_per_sample'],1), 'e2e', round(d['e2e']['value']) if d.get('e2e') else None, 'launches', d.get('gpu_launches'))
k=r.get('kernels') or {}
print({n:(round(v['alone_us'],1), round(v.get('alone_frac',0),2), v['launches_per_step']) for n,v in k.items()})
c=d.get('configs') or {}
print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in c.items() if 'value' in v})
print({k:(round(v['value']), round(v['ms_per_block'],3)) for k,v in c.get('e2e_dropin',{}).items() if isinstance(v,dict)})
cp=d.get('cpu_baseline'); print('cpu', cp['value'] if cp else None)"; }
timeout 700 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err; show $O/bench.json default
B="timeout 150 python bench.py --no-cpu --no-configs --no-e2e"
DSP_B200_FIR_T2=0 DSP_B200_FIR_STAGGER=0 $B > $O/b_old.json 2>/dev/null; show $O/b_old.json old_t4
DSP_B200_FIR_T2=0 $B --no-kernels > $O/b_s.json 2>/dev/null; show $O/b_s.json stag_t4
DSP_B200_FIR_T2=8 $B > $O/b_f8.json 2>/dev/null; show $O/b_f8.json stag_t4_f8
DSP_B200_FIR_STAGGER=0 $B --no-kernels > $O/b_f12u.json 2>/dev/null; show $O/b_f12u.json unstag_t4_f12
DSP_B200_FIR_T=6 $B --no-kernels > $O/b_t6.json 2>/dev/null; show $O/b_t6.json stag_t6_f12
for f in 12 16 8 0; do DSP_B200_FIR_T2=$f $B --block 2048 --no-kernels > $O/b2048_f$f.json 2>/dev/null; show $O/b2048_f$f.json 2048_f$f; done
DSP_B200_FIR_SINGLE_MIN=4096 $B --block 2048 --no-kernels > $O/b2048_ml.json 2>/dev/null; show $O/b2048_ml.json 2048_two_levels
$B --channels 64 --no-kernels > $O/b_c64.json 2>/dev/null; show $O/b_c64.json C64
