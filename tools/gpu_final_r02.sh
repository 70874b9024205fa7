#!/bin/bash
# final evidence with the final build: smoke, tests, bench line (+ reference arm), launch list of the bench command, and one
# full ncu capture of the three kernels of a reverse step on the light target (each ncu pass only after the same command ran
# clean without ncu)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 1 --warmup 3 > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'trunk_gemm|tailhead|enc_gemm|enc_finish|enc_split|guidance|persistent' \
    -s 9129 -c 343 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches.csv
timeout 300 python tools/profile_step.py fp16 30 > gpurun_out/profile_step.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'trunk_gemm|tailhead' -s 122 -c 3 -o gpurun_out/prof_r02_step \
    python tools/profile_step.py fp16 30 > gpurun_out/ncu_step.log 2>&1
echo "step capture rc=$?"; tail -2 gpurun_out/profile_step.log | cut -c1-300
