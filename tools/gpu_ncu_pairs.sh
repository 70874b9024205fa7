#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/perf_probe.py 5 70 20 4096 10 fp16 1 2"
timeout 300 $CMD > gpurun_out/plain_pairs.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trunk_gemm -s 20 -c 2 -o gpurun_out/prof_pairs $CMD > gpurun_out/ncu_pairs.log 2>&1
tail -3 gpurun_out/ncu_pairs.log
