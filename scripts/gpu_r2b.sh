#!/bin/bash
# round 2, visit B: co-residency of the batched MAC with the pipeline kernel (register cap x batch CTA size), the cost of
# the strided block I/O (FAKEIO timing), ncu of the slow configurations, new drop-in tests
mkdir -p gpurun_out
O=gpurun_out
Q="--no-cpu --no-e2e --no-configs"
summ() { python -c "import json,sys; d=json.load(open('$1')); print('$2', round(d['value']), round(d['ms_per_step']*1e3,1), round(d['roofline']['frac'],3), {k:(round(v['alone_us'],1), round(v.get('alone_frac',0),3)) for k,v in d['roofline'].get('kernels',{}).items()})" 2>&1 | tail -1; }
for lib in "" dsp_b200/variants/libdspb200_r128.so; do
  for bt in 256 128 64; do
    tag="lib$(basename "$lib" .so | sed 's/libdspb200//')_bt$bt"
    DSP_B200_LIB=$lib DSP_B200_FIR_BATCH_THREADS=$bt timeout 300 python bench.py $Q > $O/b_$tag.json 2> $O/b_$tag.err; summ $O/b_$tag.json $tag
  done
done
DSP_B200_FIR_PIPE_FAKEIO=1 timeout 300 python bench.py $Q > $O/b_fakeio.json 2> $O/b_fakeio.err; summ $O/b_fakeio.json fakeio
DSP_B200_FIR_PIPE_FAKEIO=1 DSP_B200_FIR_BATCH_THREADS=128 timeout 300 python bench.py $Q > $O/b_fakeio128.json 2> $O/b_fakeio128.err; summ $O/b_fakeio128.json fakeio_bt128
DSP_B200_FIR_NO_BATCH=1 timeout 300 python bench.py $Q > $O/b_nobatch.json 2> $O/b_nobatch.err; summ $O/b_nobatch.json nobatch_pf32
DSP_B200_FIR_PIPE=0 timeout 300 python bench.py $Q > $O/b_legacy.json 2> $O/b_legacy.err; summ $O/b_legacy.json legacy
DSP_B200_FIR_PIPE_EVICT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fir_pipe -s 3 -c 2 -o $O/prof_pipe_evict0 -f \
    python bench.py --steps 6 --warmup 3 $Q --no-kernels > $O/ncu_evict0.log 2>&1; tail -1 $O/ncu_evict0.log
DSP_B200_FIR_T=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fir_pipe -s 3 -c 2 -o $O/prof_pipe_t6 -f \
    python bench.py --steps 6 --warmup 3 $Q --no-kernels > $O/ncu_t6.log 2>&1; tail -1 $O/ncu_t6.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed' $O/pytest_gpu.log | tail -30
ls -la $O | tail -12
