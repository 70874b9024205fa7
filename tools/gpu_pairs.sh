#!/bin/bash
mkdir -p gpurun_out
echo "== single gemm both geometries"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "single_gemm" 2>&1 | tail -15
echo "== pair chains"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "pair_mode" 2>&1 | tail -15
for ctas in 1 2; do
  echo "== ctas $ctas"; timeout 300 python tools/perf_probe.py 5 70 20 4096 200 fp16 1 $ctas 2>&1 | tail -3
done
echo "== config3-ish"; 
timeout 300 python tools/perf_probe.py 5 1024 20 4096 20 fp16 1 1 2>&1 | tail -2
timeout 300 python tools/perf_probe.py 5 1024 20 4096 20 fp16 1 2 2>&1 | tail -2
