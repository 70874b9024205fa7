#!/bin/bash
mkdir -p gpurun_out
for v in 8 4 8 4; do
  echo "== config 2 tail_vec=$v"
  timeout 120 python tools/perf_probe.py 5 70 20 4096 1000 fp16 1 0 0 0 $v 2>&1 | tail -1
done
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
