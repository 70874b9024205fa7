#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "draws_ahead or draws_extension or remembered or shuttle" 2>&1 | tail -5
timeout 600 python tools/level1_probe.py 20 > gpurun_out/level1.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/level1.log
