#!/bin/bash
# round 2, visit O: the evidence run on the final defaults -- whole suite, smoke, full bench (both arms), launch list, ncu capture of a step
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error|exit' $O/pytest_gpu.log | tail -12
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -2 $O/smoke.log
show() { python -c "
import json,sys
try: d=json.load(open('$1'))
except Exception as e: print('$2', 'no json', e); sys.exit(0)
r=d['roofline']; k=r.get('kernels') or {}
print('$2'.ljust(20), round(d['value']), round(d['ms_per_step']*1e3,1), 'us frac', round(r['frac'],3), 'B/s', round(r['algorithmic_bytes_per_sample'],1), 'launches', d.get('gpu_launches'), {n[6:]:(round(v['alone_us'],1), round(v.get('alone_frac',0),2)) for n,v in k.items()})
c=d.get('configs') or {}
if c: print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in c.items() if 'value' in v}); print({k:(round(v['value']), round(v['ms_per_block'],3)) for k,v in c.get('e2e_dropin',{}).items() if isinstance(v,dict)})
if d.get('e2e'): print('e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'pipelined', round(d['e2e'].get('pipelined',{}).get('value',0)))
cp=d.get('cpu_baseline'); print('cpu', cp and (cp['value'], cp['cores'])); print('clocks', d.get('clocks'))"; }
timeout 800 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err; show $O/bench.json full
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-e2e --no-kernels > $O/bench20.json 2>/dev/null; show $O/bench20.json steps20
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_ref.json')); print('reference arm', d['value'], d['cpu_baseline']['cores'], d['cpu_baseline'].get('spread'))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-e2e --no-kernels > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fir_ -s 30 -c 24 -o $O/prof_fir_step -f \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-e2e --no-kernels > $O/ncu_full.log 2>&1
tail -1 $O/ncu_full.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fir_ -s 40 -c 16 -o $O/prof_fir_2048 -f \
    python bench.py --block 2048 --steps 24 --warmup 3 --no-cpu --no-configs --no-e2e --no-kernels > $O/ncu_full_2048.log 2>&1
tail -1 $O/ncu_full_2048.log
