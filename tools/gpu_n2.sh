#!/bin/bash
# 2-GPU pass: the multi-GPU tests + the config-4 bench line at N=2 (both arms launched the way the driver does)
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 rc=$?"; cut -c1-400 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "ref n2 rc=$?"; cut -c1-200 gpurun_out/bench_ref_n2.json
