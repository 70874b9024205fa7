#!/bin/bash
mkdir -p gpurun_out
CMD="python tools/perf_probe.py 5 70 20 4096 10 fp16 1 0 0 1"
timeout 200 $CMD > gpurun_out/plain_fuse.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trunk_gemm -s 13 -c 1 -o gpurun_out/prof_fuse $CMD > gpurun_out/ncu_fuse.log 2>&1
tail -2 gpurun_out/ncu_fuse.log | cut -c1-150
