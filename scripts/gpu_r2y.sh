#!/bin/bash
# round 2, visit Y: 2048-frame blocks, repeated (the figure varies from run to run): far tier capped at one CTA per SM, staggered, none
mkdir -p gpurun_out
O=gpurun_out
one() { name=$1; shift; env "$@" timeout 100 python bench.py --no-cpu --no-e2e --no-kernels --no-configs --block 2048 --steps 300 > $O/y_$name.json 2>/dev/null; python -c "
import json
d=json.load(open('$O/y_$name.json')); print('$name'.ljust(16), round(d['value']), round(d['ms_per_step']*1e3,1))"; }
for r in 1 2 3 4; do
one default_$r X=1
one cap120_$r DSP_B200_FIR_FAR_SMEM_KB=120
one cap80_$r DSP_B200_FIR_FAR_SMEM_KB=80
one twolevel_$r DSP_B200_FIR_SINGLE_MIN=4096
done
one f0_1 DSP_B200_FIR_T2=0
one f0_2 DSP_B200_FIR_T2=0
one f12_1 DSP_B200_FIR_T2=12
one f12_2 DSP_B200_FIR_T2=12
