#!/bin/bash
# compute-sanitizer memcheck over every kernel family at tiny shapes (tools/sanitize_target.py), after a plain run
mkdir -p gpurun_out
timeout 300 python tools/sanitize_target.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/sanitize_plain.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 python tools/sanitize_target.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -6 gpurun_out/sanitize_memcheck.log
