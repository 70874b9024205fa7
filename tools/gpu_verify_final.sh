#!/bin/bash
# what the driver runs at round end, with the final build: smoke, GPU tests, both bench arms
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
timeout 300 python tools/level1_probe.py 20 > gpurun_out/level1.log 2>&1; tail -3 gpurun_out/level1.log
