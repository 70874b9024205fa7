#!/bin/bash
# DRAM traffic / L2 hit rate of the GEMM at 163 840 and 256 000 rows per member (vs 20 480 at config 3)
mkdir -p gpurun_out
for n in 8192 12800; do
  timeout 600 ncu --set full --clock-control none -k regex:trunk_gemm -s 8 -c 2 -o gpurun_out/prof_rows_$n -f \
      python tools/perf_probe.py 5 $n 20 4096 2 fp16 1 > gpurun_out/ncu_rows_$n.log 2>&1
  echo "n=$n rc=$?"; tail -2 gpurun_out/ncu_rows_$n.log | cut -c1-200
done
