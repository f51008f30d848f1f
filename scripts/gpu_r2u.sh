#!/bin/bash
# round 2, visit U: ncu captures of the new K3 kernel and of a 2048-frame-block step (kept small: gpurun merges back at most 64 MiB)
mkdir -p gpurun_out
O=gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_rs_mma -s 3 -c 2 -o $O/prof_rs -f python scripts/run_resample.py > $O/ncu_rs.log 2>&1; tail -1 $O/ncu_rs.log
timeout 400 ncu --set full --clock-control none -k regex:k_fir_ -s 60 -c 16 -o $O/prof_fir_2048 -f \
    python bench.py --block 2048 --steps 24 --warmup 3 --no-cpu --no-configs --no-e2e --no-kernels > $O/ncu_full_2048.log 2>&1; tail -1 $O/ncu_full_2048.log
du -sm $O
