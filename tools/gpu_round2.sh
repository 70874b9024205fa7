#!/bin/bash
# Round-2 evidence run on one B200: GPU tests, parity report, bench, ncu launch list + one full capture.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
echo "== parity report"; timeout 600 python tools/parity_report.py > gpurun_out/parity.txt 2> gpurun_out/parity.err; echo "rc=$?"; cat gpurun_out/parity.txt | tail -40
echo "== bench (default)"; timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
if [ "$1" = "ncu" ]; then
  echo "== plain run of the profiled command"
  timeout 600 python bench.py --steps 1 --warmup 3 > gpurun_out/bench_s1.json 2> gpurun_out/bench_s1.err &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'trunk_gemm|tailhead|enc_gemm|enc_finish|enc_split|guidance' \
      -s 9048 -c 330 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_list.log 2>&1
  echo "launch list rc=$?"; tail -2 gpurun_out/ncu_list.log
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'trunk_gemm|tailhead|enc_gemm' -s 9048 -c 20 \
      -o gpurun_out/prof_r02 python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
  echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
fi
