#!/bin/bash
# Round 2 ncu evidence: launch lists (gpu__time_duration) and full captures of the dominant kernels of C3, C2 (sparse map and
# dense GEMM), C5.  Each capture after a plain run of the same command exited 0.
mkdir -p gpurun_out
P_GL="python tools/glgm06_probe.py 400 0 1"
P_C5="python tools/ensemble_probe.py 512 400"
P_BOX="python bench.py --steps 1 --warmup 1 --flow-steps 400 --no-e2e --no-cpu-baseline --no-secondary --no-riders"
run() { timeout 400 "$@" > gpurun_out/plain.log 2>&1; }
run $P_GL && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c2_r02.csv $P_GL > gpurun_out/n1.log 2>&1
run $P_GL && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"apply_r|wbox_partial|ell_left_score|chain_fast_kernel|rank_r|sort_segments|wseq_fast|finalize_r" -s 40 -c 16 -f -o gpurun_out/prof_c2_r02 $P_GL > gpurun_out/n2.log 2>&1
export REACH_GLGM06_DENSE=1
run $P_GL && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c2_dense_r02.csv $P_GL > gpurun_out/n3.log 2>&1
run $P_GL && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_dmma|score_stack" -s 14 -c 4 -f -o gpurun_out/prof_c2_dense_r02 $P_GL > gpurun_out/n4.log 2>&1
unset REACH_GLGM06_DENSE
run $P_C5 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:homog_small -s 1 -c 1 -f -o gpurun_out/prof_c5_r02 $P_C5 > gpurun_out/n5.log 2>&1
run $P_BOX && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3_r02.csv $P_BOX > gpurun_out/n6.log 2>&1
run $P_BOX && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_dmma|combine_partials|scan_block" -s 14 -c 6 -f -o gpurun_out/prof_c3_r02 $P_BOX > gpurun_out/n7.log 2>&1
tail -2 gpurun_out/n2.log gpurun_out/n4.log gpurun_out/n5.log gpurun_out/n7.log
ls -la gpurun_out/*.ncu-rep
