#!/bin/bash
# round 2, visit A: parity of the pipeline kernel, host-link probe, bench with/without the pipeline, ncu capture
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
nproc > $O/nproc.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> $O/nproc.txt
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -qi "0x10de" $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node) $(cat $d/class)"; fi; done >> $O/nproc.txt 2>&1
cat /sys/fs/cgroup/cpu.max >> $O/nproc.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_headline.py -m gpu -q --no-header -p no:cacheprovider -x > $O/pytest_headline.log 2>&1
echo "pytest headline exit $?" >> $O/pytest_headline.log; tail -15 $O/pytest_headline.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed' $O/pytest_gpu.log | tail -30
timeout 120 scripts/micro/pcie_probe > $O/pcie_probe.txt 2>&1; cat $O/pcie_probe.txt
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; cat $O/bench.json; tail -5 $O/bench.err
for v in "PIPE=0" "PIPE=1 T=6" "PIPE=1 T=8" "PIPE=1 EVICT=0" "PIPE=1 GRID=128"; do
  pipe=1; t=4; ev=1; grid=""
  for kv in $v; do case $kv in PIPE=*) pipe=${kv#PIPE=};; T=*) t=${kv#T=};; EVICT=*) ev=${kv#EVICT=};; GRID=*) grid=${kv#GRID=};; esac; done
  tag=$(echo $v | tr ' =' '__')
  DSP_B200_FIR_PIPE=$pipe DSP_B200_FIR_T=$t DSP_B200_FIR_PIPE_EVICT=$ev DSP_B200_FIR_PIPE_GRID=$grid timeout 300 python bench.py --no-cpu --no-e2e --no-configs > $O/bench_$tag.json 2> $O/bench_$tag.err
  echo "== $v"; python -c "import json,sys; d=json.load(open('$O/bench_$tag.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:(round(v['alone_us'],1), round(v.get('alone_frac',0),3)) for k,v in d['roofline'].get('kernels',{}).items()})" 2>&1 | tail -2
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs --no-kernels > $O/bench_s20.json 2>/dev/null; python -c "import json; d=json.load(open('$O/bench_s20.json')); print('steps20', d['value'], d['ms_per_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs --no-kernels > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fir_ -s 30 -c 12 -o $O/prof_fir_step -f \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-configs --no-kernels > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
ls -la $O | tail -30
