#!/bin/bash
echo "== sparse map, stream priorities"; timeout 300 python tools/glgm06_probe.py 2000 0 3 2>&1 | grep run_ms | tail -1
echo "== dense"; for K in default 8 16 24 36; do
  if [ $K = default ]; then REACH_GLGM06_DENSE=1 timeout 300 python tools/glgm06_probe.py 2000 0 3 2>&1 | grep run_ms | tail -1
  else echo "reserve $K"; REACH_GEMM_RESERVE=$K REACH_GLGM06_DENSE=1 timeout 300 python tools/glgm06_probe.py 2000 0 3 2>&1 | grep run_ms | tail -1; fi
done
