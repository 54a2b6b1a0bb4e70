#!/bin/bash
# full capture of the dominant kernel of the headline: the stacked power GEMM gemm_dmma_kernel<4,0> with the fused abs-sum epilogue
NCU="ncu --set full --clock-control none --import-source on -f"
P_BOX="python bench.py --steps 1 --warmup 1 --flow-steps 400 --no-e2e --no-cpu-baseline --no-secondary --no-riders"
timeout 400 $P_BOX > gpurun_out/plain_last.log 2>&1 && timeout 900 $NCU -k regex:"gemm_dmma" -s 90 -c 2 -o gpurun_out/prof_final_c3_power_gemm $P_BOX > gpurun_out/f12.log 2>&1
tail -n 2 gpurun_out/f12.log
python tools/ncu_summary.py gpurun_out/prof_final_c3_power_gemm.ncu-rep | cut -c1-230
ncu -i gpurun_out/prof_final_c3_power_gemm.ncu-rep --page raw --csv > gpurun_out/prof_final_c3_power_gemm_raw.csv 2>/dev/null
