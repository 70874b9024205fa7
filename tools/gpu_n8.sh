#!/bin/bash
# 8-GPU bench line only (one bounded torchrun); charged 8x box time, so nothing else runs here
mkdir -p gpurun_out
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "rc=$?"; grep '^{' gpurun_out/bench_n8.json | cut -c1-300; tail -3 gpurun_out/bench_n8.err
