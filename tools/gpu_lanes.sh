#!/bin/bash
# lanes 1 / auto / 3 at the per-rank shapes of config 4 (8, 4, 2 GPUs: 128, 256, 512 images) and at config 3
mkdir -p gpurun_out
: > gpurun_out/lanes_shards.log
for n in 128 256 512; do
  for lanes in 1 0 3; do
    timeout 200 python tools/perf_probe.py 5 $n 20 4096 $((25600 / n)) fp16 $lanes 2>&1 | tail -3 | head -1 >> gpurun_out/lanes_shards.log
  done
done
for lanes in 1 0 1 0; do
  timeout 200 python tools/perf_probe.py 5 1024 20 4096 60 fp16 $lanes 2>&1 | tail -3 | head -1 >> gpurun_out/lanes_shards.log
done
cut -c1-120 gpurun_out/lanes_shards.log
