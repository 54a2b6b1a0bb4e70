#!/bin/bash
# Round 2, call A: the full GPU parity suite (incl. the new full-length GLGM06 tests) with per-test durations, then
# compute-sanitizer memcheck / racecheck on a truncated C2 GLGM06 flowpipe and a truncated C3 BFFPSV18 flowpipe.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
tail -30 gpurun_out/pytest_gpu_a.log
for tool in memcheck racecheck; do
  timeout 700 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_c2.py 400 > gpurun_out/sanitizer_${tool}_c2_n400.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${tool}_c2_n400.log
  tail -6 gpurun_out/sanitizer_${tool}_c2_n400.log
done
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_c2.py 300 bffpsv18 > gpurun_out/sanitizer_memcheck_c3_n300.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_memcheck_c3_n300.log
tail -4 gpurun_out/sanitizer_memcheck_c3_n300.log
