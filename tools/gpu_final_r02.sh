#!/bin/bash
# final evidence with the final build: tests, bench line, launch list of the bench command
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
timeout 600 python bench.py --steps 1 --warmup 3 > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'trunk_gemm|tailhead|enc_gemm|enc_finish|enc_split|guidance|persistent' \
    -s 9129 -c 343 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
