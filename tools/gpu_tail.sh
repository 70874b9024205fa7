#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for v in 8 4 0 8 4; do
  echo "== config 2 tail_vec=$v"
  timeout 120 python tools/perf_probe.py 5 70 20 4096 1000 fp16 1 0 0 0 $v 2>&1 | tail -2
done
echo "== config 1 (K=1 N=64 D=1) tail_vec 8 / 4"
timeout 120 python tools/perf_probe.py 1 64 1 4096 1000 fp16 1 0 0 0 8 2>&1 | tail -2
timeout 120 python tools/perf_probe.py 1 64 1 4096 1000 fp16 1 0 0 0 4 2>&1 | tail -2
echo "== parity report"; timeout 600 python tools/parity_report.py > gpurun_out/parity.txt 2> gpurun_out/parity.err; cat gpurun_out/parity.txt; tail -3 gpurun_out/parity.err
