#!/bin/bash
# final evidence pass: parity tests, smoke, bench, ncu launch list + one full capture of the top kernel
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== ncu launch list"
timeout 400 python bench.py --steps 1 --warmup 3 > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'trunk_gemm|tailhead|guidance' -s 9006 -c 300 \
    --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log | cut -c1-200
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trunk_gemm -s 6010 -c 2 \
    -o gpurun_out/prof_gemm python bench.py --steps 1 --warmup 3 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-200
