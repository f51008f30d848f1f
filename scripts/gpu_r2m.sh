#!/bin/bash
# round 2, visit M: fused kernel's phase (2) as row-at-a-time load passes; tier defaults (far tier of 8 from 48 partitions on); whole suite; full bench
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -12
show() { python -c "
import json,sys
try: d=json.load(open('$1'))
except Exception as e: print('$2', 'no json', e); sys.exit(0)
r=d['roofline']; k=r.get('kernels') or {}
print('$2'.ljust(28), round(d['value']), round(d['ms_per_step']*1e3,1), 'us frac', round(r['frac'],3), 'B/s', round(r['algorithmic_bytes_per_sample'],1), {n[6:]:(round(v['alone_us'],1), round(v.get('alone_frac',0),2)) for n,v in k.items()})
c=d.get('configs') or {}
if c: print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in c.items() if 'value' in v})
if d.get('e2e'): print('e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'pipelined', round(d['e2e'].get('pipelined',{}).get('value',0)))"; }
B="timeout 120 python bench.py --no-cpu --no-configs --no-e2e --steps 300"
run() { name=$1; shift; env "$@" $B $EXTRA > $O/m_$name.json 2>/dev/null; show $O/m_$name.json $name; }
EXTRA=""
run default X=1
EXTRA="--block 2048"
run b2048_default X=1
EXTRA="--no-kernels --block 2048"
run b2048_f8s DSP_B200_FIR_STAGGER=1
run b2048_f12u DSP_B200_FIR_T2=12
EXTRA="--no-kernels --channels 64"
run c64_default X=1
timeout 500 python bench.py --no-cpu --no-kernels > $O/bench_m.json 2> $O/bench_m.err; show $O/bench_m.json full_default
