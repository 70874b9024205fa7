#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for ctas in 1 2 0; do
  echo "== ctas $ctas"; timeout 300 python tools/perf_probe.py 5 70 20 4096 200 fp16 1 $ctas 1 2>&1 | tail -2 | cut -c1-330
done
echo "== config3-ish auto"; timeout 300 python tools/perf_probe.py 5 1024 20 4096 20 fp16 1 0 1 2>&1 | tail -2 | cut -c1-330
echo "== single member B=64"; timeout 300 python tools/perf_probe.py 1 64 1 4096 200 fp16 1 0 1 2>&1 | tail -2 | cut -c1-330
