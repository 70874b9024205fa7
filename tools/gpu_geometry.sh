#!/bin/bash
# auto geometry (makespan-aware) vs forced single / pairs where the choice changed: one member x 1400 rows (draws-ahead),
# K=5 x 180 rows (config 2 strong-scaled over 8 GPUs), and the unchanged config 2 as control
mkdir -p gpurun_out
: > gpurun_out/geometry.log
run() { timeout 200 python tools/perf_probe.py $1 $2 $3 4096 $4 fp16 1 $5 2>&1 | tail -3 | head -1 | cut -c1-150 >> gpurun_out/geometry.log; }
for c in 0 1 2 0 1 2; do run 1 70 20 400 $c; done
for c in 0 1 2 0 1 2; do run 5 9 20 600 $c; done
for c in 0 1 2; do run 5 70 20 300 $c; done
for c in 0 1 2; do run 2 70 20 400 $c; done
for c in 0 1 2; do run 1 140 20 300 $c; done
sed -e 's/ host enqueue.*launches/ launches/' gpurun_out/geometry.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "packed or pair or properties or shape_matrix" 2>&1 | tail -3
