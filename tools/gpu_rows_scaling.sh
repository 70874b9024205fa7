#!/bin/bash
# per-kernel times vs rows per member (why do the 0.5M+-chain sweep points run at ~1050 TFLOP/s against 1230 at config 3?)
mkdir -p gpurun_out
: > gpurun_out/rows_scaling.log
for n in 1024 4096 8192 12800; do
  timeout 300 python tools/perf_probe.py 5 $n 20 4096 $((20480 / n + 6)) fp16 1 2>&1 | tail -2 | cut -c1-330 >> gpurun_out/rows_scaling.log
done
for o in 1 2; do
  timeout 300 python tools/perf_probe.py 5 8192 20 4096 8 fp16 1 0 0 0 0 $o 2>&1 | tail -2 | cut -c1-330 >> gpurun_out/rows_scaling.log
done
cat gpurun_out/rows_scaling.log
