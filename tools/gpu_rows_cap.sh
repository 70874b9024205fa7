#!/bin/bash
# rows-per-call cap of NestedEnsemble: 262144 (round-1 default) vs 65536 (new default) vs 32768 at the large sweep points
mkdir -p gpurun_out
: > gpurun_out/rows_cap.log
export SWEEP_BUDGET=4e7 SWEEP_REPS=2 SWEEP_CPU=0
for cap in 262144 65536 32768 262144 65536; do
  echo "cap $cap" >> gpurun_out/rows_cap.log
  LADINE_MAX_ROWS=$cap timeout 300 python tools/sweep_dist.py points 1000,10,16384 1000,100,1024 1000,1000,1024 2>&1 | grep '^{' | cut -c1-200 >> gpurun_out/rows_cap.log
done
cat gpurun_out/rows_cap.log
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x -k "tiling or scattered or ensemble" 2>&1 | tail -3
