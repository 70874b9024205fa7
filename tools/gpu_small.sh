#!/bin/bash
mkdir -p gpurun_out
echo "== gemm layer + invariance tests"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "single_gemm_layer or pair_mode_equals" 2>&1 | tail -5
for c in 1 3 0; do
  echo "== config 1 (K=1 N=64 D=1) ctas=$c"
  timeout 120 python tools/perf_probe.py 1 64 1 4096 1000 fp16 1 $c 0 0 0 2>&1 | tail -2
done
for c in 1 3; do
  echo "== K=1 N=70 D=4 ctas=$c"; timeout 120 python tools/perf_probe.py 1 70 4 4096 1000 fp16 1 $c 0 0 0 2>&1 | tail -1
  echo "== K=5 N=64 D=1 ctas=$c"; timeout 120 python tools/perf_probe.py 5 64 1 4096 1000 fp16 1 $c 0 0 0 2>&1 | tail -1
  echo "== K=5 N=70 D=2 (10 wide tiles/N tile... 160 wide tiles) ctas=$c"; timeout 120 python tools/perf_probe.py 5 70 2 4096 1000 fp16 1 $c 0 0 0 2>&1 | tail -1
  echo "== config 2 ctas=$c"; timeout 120 python tools/perf_probe.py 5 70 20 4096 1000 fp16 1 $c 0 0 0 2>&1 | tail -1
done
echo "== full gpu suite"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
