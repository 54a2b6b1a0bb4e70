#!/bin/bash
# Round 2, final code: `ncu --set full` of one steady-state instance of every kernel family on the path, plus launch lists
# (gpu__time_duration) of the C3 / C2 / C2-dense probes.  Every capture follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
LL="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
P_GL="python tools/glgm06_probe.py 400 3 1"
P_BOX="python bench.py --steps 1 --warmup 1 --flow-steps 400 --no-e2e --no-cpu-baseline --no-secondary --no-riders"
P_C5="python tools/ensemble_probe.py 512 400"
P_BIG="python tools/glgm06_large_sharded.py 512 140 10"
P_LZ="python tools/lazy_probe.py 32 200 0"
P_PR="python tools/project_probe.py 400"
ok() { timeout 400 "$@" > gpurun_out/plain_last.log 2>&1; }
# --- GLGM06 C2, sparse map: every kernel of a steady-state 64-step batch
ok $P_GL && timeout 600 $LL -c 900 --log-file gpurun_out/launches_c2_final.csv $P_GL > gpurun_out/f1.log 2>&1
ok $P_GL && timeout 900 $NCU -k regex:"ell_left_score|sort_segments|chain_decide|chain_fast|wbox_partial|wseq|rank_r|apply_r|finalize_r|materialize" -s 104 -c 16 -o gpurun_out/prof_final_c2 $P_GL > gpurun_out/f2.log 2>&1
# --- GLGM06 C2, dense map
export REACH_GLGM06_DENSE=1
ok $P_GL && timeout 600 $LL -c 900 --log-file gpurun_out/launches_c2_dense_final.csv $P_GL > gpurun_out/f3.log 2>&1
ok $P_GL && timeout 900 $NCU -k regex:"gemm_dmma|score_stack" -s 16 -c 4 -o gpurun_out/prof_final_c2_dense $P_GL > gpurun_out/f4.log 2>&1
unset REACH_GLGM06_DENSE
# --- dense GLGM06 at n = 1024 (GEMM in situ) and the device zonotope discretisation
ok $P_BIG && timeout 900 $NCU -k regex:"gemm_dmma" -s 24 -c 2 -o gpurun_out/prof_final_n1024_gemm $P_BIG > gpurun_out/f5.log 2>&1
ok $P_BIG && timeout 900 $NCU -k regex:"zono|box_matvec|omega0|sih|col_abs|scale|axpby|hull|compact" -c 16 -o gpurun_out/prof_final_disc $P_BIG > gpurun_out/f6.log 2>&1
# --- BFFPSV18 C3
ok $P_BOX && timeout 600 $LL -c 400 --log-file gpurun_out/launches_c3_final.csv $P_BOX > gpurun_out/f7.log 2>&1
ok $P_BOX && timeout 900 $NCU -k regex:"gemm_dmma|scan_block|combine|gather_rows" -s 14 -c 6 -o gpurun_out/prof_final_c3 $P_BOX > gpurun_out/f8.log 2>&1
# --- C5 ensemble, lazy exponential, projection
ok $P_C5 && timeout 600 $NCU -k regex:homog_small -s 1 -c 1 -o gpurun_out/prof_final_c5 $P_C5 > gpurun_out/f9.log 2>&1
export REACH_LAZY_NO_GRAPH=1
ok $P_LZ && timeout 600 $NCU -k regex:"spmm_taylor|lazy_partial|lazy_finish" -s 60 -c 6 -o gpurun_out/prof_final_lazy $P_LZ > gpurun_out/f10.log 2>&1
unset REACH_LAZY_NO_GRAPH
ok $P_PR && timeout 600 $NCU -k regex:"project_" -c 4 -o gpurun_out/prof_final_project $P_PR > gpurun_out/f11.log 2>&1
for f in gpurun_out/f*.log; do echo "== $f"; tail -1 $f; done
ls -la gpurun_out/prof_final_*.ncu-rep gpurun_out/launches_*_final.csv
# gpurun brings back at most 64 MiB: summarise on the box, keep the raw pages of the two largest reports as CSV instead
python tools/ncu_summary.py gpurun_out/prof_final_c2.ncu-rep gpurun_out/prof_final_c2_dense.ncu-rep gpurun_out/prof_final_n1024_gemm.ncu-rep \
  gpurun_out/prof_final_c3.ncu-rep gpurun_out/prof_final_c5.ncu-rep gpurun_out/prof_final_disc.ncu-rep gpurun_out/prof_final_lazy.ncu-rep \
  gpurun_out/prof_final_project.ncu-rep > gpurun_out/ncu_all_kernels_final_summary.txt 2>&1
for r in c2 disc; do ncu -i gpurun_out/prof_final_$r.ncu-rep --page raw --csv > gpurun_out/prof_final_${r}_raw.csv 2>/dev/null; rm -f gpurun_out/prof_final_$r.ncu-rep; done
du -sh gpurun_out
