#!/bin/bash
# ncu --set full capture of one kernel (regex $1) of the bench step; report -> gpurun_out/prof_$2.ncu-rep
pat=$1; name=$2; shift 2
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$pat -s 6 -c 2 -o gpurun_out/prof_$name -f \
    python bench.py --steps 12 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/ncu_$name.log 2>&1
tail -2 gpurun_out/ncu_$name.log
