#!/bin/bash
for ctas in 1 2; do for dbg in 0 1 2; do
  echo "== ctas $ctas dbg $dbg"; LADINE_GEMM_DBG=$dbg timeout 300 python tools/perf_probe.py 5 70 20 4096 200 fp16 1 $ctas 1 2>&1 | tail -1 | cut -c1-200
done; done
