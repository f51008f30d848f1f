#!/bin/bash
# round 2, visit V (4 GPUs): does the host-call path scale over the box's NUMA nodes?
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
cat /sys/fs/cgroup/cpu.max 2>/dev/null
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 4 --steps 200 --warmup 5 --no-kernels > $O/bench_n4.json 2> $O/bench_n4.err; echo "n4 exit $?"; tail -2 $O/bench_n4.err
python -c "
import json
d=json.loads(open('$O/bench_n4.json').read().strip().splitlines()[-1])
print('N=4', round(d['value']), round(d['ms_per_step']*1e3,1), 'us', 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d['e2e'].get('numa'), 'pipelined', round(d['e2e']['pipelined']['value']))"
