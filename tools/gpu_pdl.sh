#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for pdl in 0 1; do
  echo "== pdl $pdl"; timeout 300 python tools/perf_probe.py 5 70 20 4096 200 fp16 1 1 $pdl 2>&1 | tail -3 | cut -c1-400
done
echo "== pdl 1 lanes 2"; timeout 300 python tools/perf_probe.py 5 70 20 4096 200 fp16 2 1 1 2>&1 | tail -3 | cut -c1-300
