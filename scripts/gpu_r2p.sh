#!/bin/bash
# round 2, visit P: the evidence artefacts (bench line, launch list, one ncu capture of four steps) -- kept under gpurun's 64 MiB
mkdir -p gpurun_out
O=gpurun_out
show() { python -c "
import json,sys
try: d=json.load(open('$1'))
except Exception as e: print('$2', 'no json', e); sys.exit(0)
r=d['roofline']; k=r.get('kernels') or {}
print('$2'.ljust(20), round(d['value']), round(d['ms_per_step']*1e3,1), 'us frac', round(r['frac'],3), 'B/s', round(r['algorithmic_bytes_per_sample'],1), 'launches', d.get('gpu_launches'), {n[6:]:(round(v['alone_us'],1), round(v.get('alone_frac',0),2)) for n,v in k.items()})
c=d.get('configs') or {}
if c: print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in c.items() if 'value' in v}); print({k:(round(v['value']), round(v['ms_per_block'],3)) for k,v in c.get('e2e_dropin',{}).items() if isinstance(v,dict)})
if d.get('e2e'): print('e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'pipelined', round(d['e2e'].get('pipelined',{}).get('value',0)))
cp=d.get('cpu_baseline'); print('cpu', cp and (cp['value'], cp['cores']))"; }
timeout 800 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err; show $O/bench.json full
DSP_B200_FIR_NO_DIRECT=1 timeout 150 python bench.py --no-cpu --no-configs --no-e2e --steps 300 > $O/p_nodirect.json 2>/dev/null; show $O/p_nodirect.json staged_io
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_ref.json')); print('reference arm', d['value'], d['cpu_baseline']['cores'], d['cpu_baseline'].get('spread'))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-e2e --no-kernels > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fir_ -s 30 -c 12 -o $O/prof_fir_step -f \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-e2e --no-kernels > $O/ncu_full.log 2>&1
tail -1 $O/ncu_full.log
du -sm $O; ls -la $O | head -20
if [ $(du -sm $O | cut -f1) -gt 60 ]; then echo "too big: dropping the ncu report"; ncu -i $O/prof_fir_step.ncu-rep --page raw --csv > $O/prof_fir_step_raw.csv 2>/dev/null; rm -f $O/prof_fir_step.ncu-rep; fi
