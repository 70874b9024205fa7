#!/bin/bash
# first GPU bring-up: each stage in its own process so a trapped kernel cannot poison the next
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
echo "== resident tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "small or c10 or c3 or cosine or ensemble_k3 or single_step or error" 2>&1 | tail -25
echo "== single gemm"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "single_gemm" 2>&1 | tail -40
echo "== tensor chains"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "tc_ or philox or partition or ensemble_tc" 2>&1 | tail -40
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8
echo "== perf"; timeout 300 python tools/perf_probe.py 5 70 20 4096 50 fp16 2>&1 | tail -5
timeout 300 python tools/perf_probe.py 1 64 1 128 1000 fp32 2>&1 | tail -3
