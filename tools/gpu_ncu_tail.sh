#!/bin/bash
# one full ncu capture of the tail/head kernel inside the bench workload (after the same command ran clean without ncu)
mkdir -p gpurun_out
timeout 400 python bench.py --steps 1 --warmup 3 > gpurun_out/plain_tail.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tailhead -s 3010 -c 2 \
    -o gpurun_out/prof_tail python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_tail.log 2>&1
tail -2 gpurun_out/ncu_tail.log | cut -c1-200
