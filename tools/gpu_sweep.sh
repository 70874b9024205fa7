#!/bin/bash
# every configuration in its own process under a tight timeout
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
one() { timeout 150 python tools/sweep.py one "$@" >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err || echo "{\"failed\": \"$*\", \"rc\": $?}" >> gpurun_out/sweep.jsonl; }
one 1000 1 64 1
one 1000 20 70 5
one 1000 100 70 5
one 1000 20 1024 5
for T in 100 1000; do for D in 10 100 1000; do for N in 64 1024 16384; do
  chains=$((N*D*5)); if [ $chains -le 100000000 ]; then one $T $D $N 5; fi
done; done; done
cat gpurun_out/sweep.jsonl
