#!/bin/bash
# config-5 sweep at 1 GPU with the 32768-row cap (sampler only, through sample_ensemble; CPU column at the small end)
mkdir -p gpurun_out
SWEEP_BUDGET=3e8 SWEEP_REPS=2 timeout 1200 python tools/sweep_dist.py > gpurun_out/sweep_n1.jsonl 2> gpurun_out/sweep_n1.err
echo "rc=$?"; cut -c1-140 gpurun_out/sweep_n1.jsonl; tail -2 gpurun_out/sweep_n1.err
