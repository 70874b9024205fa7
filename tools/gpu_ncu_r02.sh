#!/bin/bash
# ncu evidence, round 2.  (1) launch list of the bench command itself; (2) full captures on the light target.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'trunk_gemm|tailhead|enc_gemm|enc_finish|enc_split|guidance' \
    -s 9129 -c 343 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_list.log; wc -l gpurun_out/launches.csv
timeout 300 python tools/profile_step.py fp16 30 > gpurun_out/profile_step.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:enc_gemm -s 6 -c 3 -o gpurun_out/prof_r02_enc \
    python tools/profile_step.py fp16 30 > gpurun_out/ncu_enc.log 2>&1
echo "enc capture rc=$?"; cat gpurun_out/profile_step.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'trunk_gemm|tailhead' -s 122 -c 3 -o gpurun_out/prof_r02_step \
    python tools/profile_step.py fp16 30 > gpurun_out/ncu_step.log 2>&1
echo "step capture rc=$?"
timeout 300 python tools/profile_step.py fp32x 10 > gpurun_out/profile_step_fp32x.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'trunk_split' -s 32 -c 2 -o gpurun_out/prof_r02_split \
    python tools/profile_step.py fp32x 10 > gpurun_out/ncu_split.log 2>&1
echo "split capture rc=$?"; cat gpurun_out/profile_step_fp32x.log
ls -la gpurun_out/
