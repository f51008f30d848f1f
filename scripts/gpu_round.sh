#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench, launch list + one full ncu capture.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [quick|full]
mode=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E 'FAILED|ERROR|passed|failed' gpurun_out/pytest_gpu.log | tail -40
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$mode" = "full" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
  # every k_fir_* launch of 8 steps (8 fused level-0, 8 per-block MACs, 2 batched MACs): per-kernel DRAM traffic for traffic.json
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fir_ -s 45 -c 18 -o gpurun_out/prof_fir_step -f \
      python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bq_cascade -s 3 -c 1 -o gpurun_out/prof_bq -f \
      python scripts/run_biquad.py > gpurun_out/ncu_bq.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rs_mma -s 3 -c 1 -o gpurun_out/prof_rs -f \
      python scripts/run_resample.py 1024 > gpurun_out/ncu_rs.log 2>&1
  timeout 300 python scripts/bench_kernels.py > gpurun_out/kernels.json 2> gpurun_out/kernels.err
  ls -la gpurun_out
fi
