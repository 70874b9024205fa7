#!/bin/bash
# rows-per-call cap, second pass: below 32768, including config 3 itself (20480 rows per member)
mkdir -p gpurun_out
: > gpurun_out/rows_cap2.log
export SWEEP_BUDGET=4e7 SWEEP_REPS=2 SWEEP_CPU=0
for cap in 32768 16384 10240 5120 32768 10240; do
  echo "cap $cap" >> gpurun_out/rows_cap2.log
  LADINE_MAX_ROWS=$cap timeout 300 python tools/sweep_dist.py points 1000,20,1024 1000,10,16384 1000,1000,1024 2>&1 | grep '^{' | cut -c1-175 >> gpurun_out/rows_cap2.log
done
cat gpurun_out/rows_cap2.log
