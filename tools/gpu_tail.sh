#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
CMD="python tools/perf_probe.py 5 70 20 4096 200 fp16 1 1 1"
timeout 300 $CMD 2>&1 | tail -2 | cut -c1-330
CMD2="python tools/perf_probe.py 5 70 20 4096 8 fp16 1 1 1"
timeout 300 $CMD2 > gpurun_out/plain_tail.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tailhead -s 8 -c 2 -o gpurun_out/prof_tail $CMD2 > gpurun_out/ncu_tail.log 2>&1
tail -2 gpurun_out/ncu_tail.log
