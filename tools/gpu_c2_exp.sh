#!/bin/bash
# C2 experiments: the probe under different SM-sharing policies ("exclusive reserve" pairs in $@; reserve "-" = heuristic).
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  echo "=== REACH_CHAIN_EXCLUSIVE=$1 REACH_GEMM_RESERVE=$2 DETAIL=$3" >> gpurun_out/c2_exp.log
  if [ "$2" = "-" ]; then unset REACH_GEMM_RESERVE; else export REACH_GEMM_RESERVE=$2; fi
  if [ "$3" = "1" ]; then export REACH_TIMING_DETAIL=1; else unset REACH_TIMING_DETAIL; fi
  REACH_CHAIN_EXCLUSIVE=$1 timeout 300 python tools/glgm06_probe.py 2000 0 4 >> gpurun_out/c2_exp.log 2>&1
done
grep -E "===|steps_per_s|detail" gpurun_out/c2_exp.log | cut -c1-330
