#!/bin/bash
# round 2, visit AE: the headline in fresh processes, block cache on / off (one full run read 104.7 us)
mkdir -p gpurun_out
O=gpurun_out
one() { name=$1; shift; env "$@" timeout 100 python bench.py --no-cpu --no-e2e --no-kernels --no-configs > $O/ae_$name.json 2>/dev/null; python -c "
import json
d=json.load(open('$O/ae_$name.json')); print('$name'.ljust(12), round(d['value']), round(d['ms_per_step']*1e3,1))"; }
one pool_1 X=1
one nopool_1 DSP_B200_POOL_MB=0
one pool_2 X=1
one nopool_2 DSP_B200_POOL_MB=0
one pool_3 X=1
one pool_4 X=1
