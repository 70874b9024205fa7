#!/bin/bash
mkdir -p gpurun_out
echo "== invariance test"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pair_mode_equals" 2>&1 | tail -3
for o in 1 2 1 2; do
  echo "== config 3 (K=5 N=1024 D=20) order=$o"
  timeout 150 python tools/perf_probe.py 5 1024 20 4096 60 fp16 1 0 0 0 0 $o 2>&1 | tail -2
done
for o in 1 2; do
  echo "== K=5 N=256 D=20 (25.6k chains, 42 MB act/member) order=$o"
  timeout 150 python tools/perf_probe.py 5 256 20 4096 200 fp16 1 0 0 0 0 $o 2>&1 | tail -1
  echo "== K=5 N=512 D=20 (51k chains, 84 MB act/member) order=$o"
  timeout 150 python tools/perf_probe.py 5 512 20 4096 100 fp16 1 0 0 0 0 $o 2>&1 | tail -1
  echo "== config 2 order=$o"
  timeout 150 python tools/perf_probe.py 5 70 20 4096 1000 fp16 1 0 0 0 0 $o 2>&1 | tail -1
done
