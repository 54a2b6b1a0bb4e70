#!/bin/bash
# Final validation of a round on one B200: GPU tests, smoke, bench (ours + reference arm), then ncu evidence of the C2 kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1
P_GL="python tools/glgm06_probe.py 400 0 1"
timeout 200 $P_GL > gpurun_out/plain_gl.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_glgm06_v4.csv $P_GL > gpurun_out/ncu_gl_l.log 2>&1
timeout 200 $P_GL > gpurun_out/plain_gl2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ell_left|chain_kernel|apply_r|sort_segments" -s 30 -c 8 -f -o gpurun_out/prof_glgm06_v4 $P_GL > gpurun_out/ncu_gl_f.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; tail -2 gpurun_out/bench.log | cut -c1-400; tail -1 gpurun_out/bench_ref.log | cut -c1-300; tail -2 gpurun_out/ncu_gl_f.log
