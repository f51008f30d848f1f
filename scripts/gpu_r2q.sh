#!/bin/bash
# round 2, visit Q (2 GPUs): the bench under torchrun, both arms, as the driver launches them
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "n2 exit $?"; tail -2 $O/bench_n2.err
python -c "
import json
d=json.loads(open('$O/bench_n2.json').read().strip().splitlines()[-1])
print('N=2', round(d['value']), round(d['ms_per_step']*1e3,1), 'us', 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), d['e2e'].get('numa'), 'launches', d.get('gpu_launches'), d['config'])"
timeout 300 $TR --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 exit $?"
python -c "
import json
d=json.loads(open('$O/bench_ref_n2.json').read().strip().splitlines()[-1]); print('reference N=2', d['value'], d.get('n_gpus'), d['cpu_baseline']['cores'])"
timeout 200 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu --no-configs --no-kernels > $O/bench_n1.json 2>/dev/null
python -c "
import json
d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1])
print('N=1', round(d['value']), round(d['ms_per_step']*1e3,1), 'us', 'e2e', round(d['e2e']['value']))"
