#!/bin/bash
# round 2, visit D: whole suite on the default (three-kernel) path + opt-in pipeline tests, full bench, host-call slab sweep, K1 split, ncu
mkdir -p gpurun_out
O=gpurun_out
Q="--no-cpu --no-configs --no-kernels --steps 200"
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed|Error' $O/pytest_gpu.log | tail -20
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('H', round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), 'pipelined', round(d['e2e']['pipelined']['value'])); print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1), round(v.get('roofline',{}).get('frac',0),3)) for k,v in d['configs'].items() if 'value' in v}); print({k:(round(v['value']), round(v['ms_per_block'],3)) for k,v in d['configs'].get('e2e_dropin',{}).items() if isinstance(v,dict)}); print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
for sl in "1 1" "2 1" "2 2" "8 2"; do set -- $sl
  timeout 200 python bench.py $Q --e2e-slabs $1 --e2e-pipe-slabs $2 > $O/e2e_$1_$2.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/e2e_$1_$2.json')); print('slabs $1 / pipe $2: sync', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'pipelined', round(d['e2e']['pipelined']['value']), round(d['e2e']['pipelined']['ms_per_step'],3))"
done
DSP_B200_BQ_SPLIT=0 timeout 120 python scripts/bench_biquad.py > $O/bq_split0.txt 2>&1; tail -2 $O/bq_split0.txt
DSP_B200_BQ_SPLIT=1 timeout 120 python scripts/bench_biquad.py > $O/bq_split1.txt 2>&1; tail -2 $O/bq_split1.txt
DSP_B200_FIR_PIPE=1 timeout 200 python bench.py --no-cpu --no-configs --no-e2e > $O/bench_pipe.json 2>/dev/null; python -c "
import json; d=json.load(open('$O/bench_pipe.json')); print('pipe', round(d['value']), d['ms_per_step'], round(d['roofline']['frac'],3), {k:(round(v['alone_us'],1), round(v.get('alone_frac',0),3)) for k,v in d['roofline'].get('kernels',{}).items()})"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-e2e --no-kernels > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fir_ -s 30 -c 18 -o $O/prof_fir_step -f \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --no-e2e --no-kernels > $O/ncu_full.log 2>&1
tail -1 $O/ncu_full.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_bq_cascade -s 3 -c 2 -o $O/prof_bq -f \
    python scripts/run_biquad.py > $O/ncu_bq.log 2>&1; tail -1 $O/ncu_bq.log
