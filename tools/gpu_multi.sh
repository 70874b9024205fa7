#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "== pytest multi + full"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "== bench 1 gpu"; timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-900 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
echo "== bench 2 gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cut -c1-900 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
