#!/bin/bash
# round 2, visit C: the one-kernel-per-block pipeline (batched tail inside), stats, variants, full bench, host-link probe
mkdir -p gpurun_out
O=gpurun_out
Q="--no-cpu --no-e2e --no-configs"
summ() { python -c "import json,sys; d=json.load(open('$1')); print('$2', round(d['value']), round(d['ms_per_step']*1e3,1), round(d['roofline']['frac'],3), {k:(round(v['alone_us'],1), round(v.get('alone_frac',0),3)) for k,v in d['roofline'].get('kernels',{}).items()})" 2>&1 | tail -1; }
timeout 900 python -m pytest tests/test_gpu_headline.py -m gpu -q --no-header -p no:cacheprovider -x > $O/pytest_headline.log 2>&1
echo "pytest headline exit $?" >> $O/pytest_headline.log; tail -12 $O/pytest_headline.log
for i in 1 2 3; do timeout 300 python -m pytest "tests/test_gpu_headline.py::test_pipe_more_channels_than_sms" -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3; done
timeout 300 python bench.py $Q > $O/c_mega.json 2> $O/c_mega.err; summ $O/c_mega.json mega; tail -3 $O/c_mega.err
timeout 300 python scripts/pipe_stats.py > $O/pipe_stats.txt 2>&1; cat $O/pipe_stats.txt
DSP_B200_FIR_PIPE_NOITEMS=1 timeout 300 python bench.py $Q > $O/c_noitems.json 2> $O/c_noitems.err; summ $O/c_noitems.json noitems
DSP_B200_FIR_PIPE_EVICT=0 timeout 300 python bench.py $Q > $O/c_evict0.json 2> $O/c_evict0.err; summ $O/c_evict0.json evict0
DSP_B200_FIR_PIPE_EVICT=0 timeout 300 python scripts/pipe_stats.py > $O/pipe_stats_evict0.txt 2>&1; cat $O/pipe_stats_evict0.txt
DSP_B200_FIR_PIPE_FAKEIO=1 timeout 300 python bench.py $Q > $O/c_fakeio.json 2> $O/c_fakeio.err; summ $O/c_fakeio.json fakeio
DSP_B200_FIR_PIPE_GRID=128 timeout 300 python bench.py $Q > $O/c_grid128.json 2> $O/c_grid128.err; summ $O/c_grid128.json grid128
DSP_B200_FIR_PIPE=0 timeout 300 python bench.py $Q > $O/c_legacy.json 2> $O/c_legacy.err; summ $O/c_legacy.json legacy
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed' $O/pytest_gpu.log | tail -30
timeout 120 scripts/micro/pcie_probe > $O/pcie_probe.txt 2>&1; tail -24 $O/pcie_probe.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['pipelined']['value']); print({k:(round(v.get('value',0)), round(v.get('ms_per_block',0)*1e3,1)) for k,v in d['configs'].items() if 'value' in v}); print(d['configs'].get('e2e_dropin')); print(d['cpu_baseline']['value'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 $Q --no-kernels > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fir_pipe -s 8 -c 3 -o $O/prof_fir_pipe -f \
    python bench.py --steps 12 --warmup 3 $Q --no-kernels > $O/ncu_full.log 2>&1
tail -1 $O/ncu_full.log
ls -la $O | tail -8
