#!/bin/bash
# final round-2 verification on one B200: smoke, GPU tests, parity report, bench, ncu of the persistent kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
echo "== parity report"; timeout 600 python tools/parity_report.py > gpurun_out/parity.txt 2> gpurun_out/parity.err; echo "rc=$?"; tail -12 gpurun_out/parity.txt
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cat gpurun_out/bench_n1.json | cut -c1-1500; tail -3 gpurun_out/bench_n1.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-600
echo "== persistent kernel ncu"
timeout 300 python tools/profile_small.py 64 fp16 > gpurun_out/profile_small.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:persistent_chain -s 2 -c 1 -o gpurun_out/prof_r02_persist \
    python tools/profile_small.py 64 fp16 > gpurun_out/ncu_persist.log 2>&1
echo "persist capture rc=$?"; head -4 gpurun_out/profile_small.log; ls -la gpurun_out/*.ncu-rep
