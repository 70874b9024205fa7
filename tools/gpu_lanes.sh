#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for lanes in 1 2 3; do
  echo "== lanes $lanes"; timeout 300 python tools/perf_probe.py 5 70 20 4096 200 fp16 $lanes 2>&1 | tail -3
done
echo "== config3-ish lanes 2"; timeout 300 python tools/perf_probe.py 5 1024 20 4096 20 fp16 2 2>&1 | tail -2
echo "== single member lanes"; timeout 300 python tools/perf_probe.py 1 64 1 4096 200 fp16 2 2>&1 | tail -2
