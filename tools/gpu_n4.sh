#!/bin/bash
# 4-GPU bench line (config 4 at N=4), bounded
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "rc=$?"; grep '^{' gpurun_out/bench_n4.json | cut -c1-300; tail -3 gpurun_out/bench_n4.err
