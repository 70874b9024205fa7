#!/bin/bash
echo "== bitwise + pair tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "bitwise or pair_mode or shape_matrix or tc_f256" 2>&1 | tail -5
echo "== full"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for fuse in 0 1; do
  echo "== fuse $fuse"; timeout 120 python tools/perf_probe.py 5 70 20 4096 200 fp16 1 0 0 $fuse 2>&1 | tail -2 | cut -c1-330
done
