#!/bin/bash
# the two captures of gpu_s2_ncu_all.sh that landed in the discretisation's small GEMMs: match the full template instance
NCU="ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
P_BOX="python bench.py --steps 1 --warmup 1 --flow-steps 400 --no-e2e --no-cpu-baseline --no-secondary --no-riders"
P_BIG="python tools/glgm06_large_sharded.py 512 140 10"
timeout 400 $P_BOX > gpurun_out/plain_last.log 2>&1 && timeout 900 $NCU -k regex:"gemm_dmma_kernel<4, false>|gemm_dmma_kernel<4, 0>|scan_block_kernel" -s 12 -c 6 -o gpurun_out/prof_final_c3 $P_BOX > gpurun_out/f8.log 2>&1
timeout 400 $P_BIG > gpurun_out/plain_last.log 2>&1 && timeout 900 $NCU -k regex:"gemm_dmma_kernel" -s 120 -c 6 -o gpurun_out/prof_final_n1024_gemm $P_BIG > gpurun_out/f5.log 2>&1
tail -2 gpurun_out/f8.log gpurun_out/f5.log
python tools/ncu_summary.py gpurun_out/prof_final_c3.ncu-rep gpurun_out/prof_final_n1024_gemm.ncu-rep > gpurun_out/ncu_final_fix_summary.txt 2>&1
cat gpurun_out/ncu_final_fix_summary.txt | cut -c1-220
