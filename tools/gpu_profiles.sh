#!/bin/bash
# One gpurun call: ncu evidence for the current kernels (each capture after a plain run of the same command exited 0).
mkdir -p gpurun_out
P_GL="python tools/glgm06_probe.py 400 0 1"
P_LZ="python tools/lazy_probe.py 32 60 0"
P_LZ512="python tools/lazy_probe.py 512 30 0"
P_BOX="python bench.py --steps 1 --warmup 1 --flow-steps 400 --no-e2e --no-cpu-baseline --no-secondary"
timeout 300 $P_GL > gpurun_out/plain_gl.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_glgm06_v3.csv $P_GL > gpurun_out/ncu_gl_l.log 2>&1
timeout 300 $P_GL > gpurun_out/plain_gl2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"apply_r|wbox_partial|gemm_dmma|chain_kernel|rank_r|score_stack|sort_segments" -s 60 -c 14 -f -o gpurun_out/prof_glgm06_v3 $P_GL > gpurun_out/ncu_gl_f.log 2>&1
timeout 300 $P_LZ > gpurun_out/plain_lz.log 2>&1 && \
REACH_LAZY_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_lazy.csv $P_LZ > gpurun_out/ncu_lz_l.log 2>&1
timeout 300 $P_LZ > gpurun_out/plain_lz2.log 2>&1 && \
REACH_LAZY_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_taylor|lazy_partial|lazy_finish" -s 200 -c 6 -f -o gpurun_out/prof_lazy $P_LZ > gpurun_out/ncu_lz_f.log 2>&1
timeout 300 $P_LZ512 > gpurun_out/plain_lz3.log 2>&1 && \
REACH_LAZY_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_taylor" -s 100 -c 2 -f -o gpurun_out/prof_lazy512 $P_LZ512 > gpurun_out/ncu_lz_f512.log 2>&1
timeout 300 python tools/project_probe.py > gpurun_out/plain_pr.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"project_" -c 2 -f -o gpurun_out/prof_project python tools/project_probe.py > gpurun_out/ncu_pr.log 2>&1
timeout 300 $P_BOX > gpurun_out/plain_box.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_dmma -s 12 -c 2 -f -o gpurun_out/prof_gemm_v3 $P_BOX > gpurun_out/ncu_box_f.log 2>&1
tail -2 gpurun_out/ncu_gl_f.log gpurun_out/ncu_lz_f.log gpurun_out/ncu_pr.log gpurun_out/ncu_box_f.log; cat gpurun_out/plain_pr.log | tail -3
