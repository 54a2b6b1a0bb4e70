#!/bin/bash
echo "== dense"; for K in default 2 4 8 12 16 24; do
  if [ $K = default ]; then REACH_GLGM06_DENSE=1 timeout 300 python tools/glgm06_probe.py 2000 0 3 2>&1 | grep run_ms | tail -1 | cut -c1-230
  else echo "reserve $K"; REACH_GEMM_RESERVE=$K REACH_GLGM06_DENSE=1 timeout 300 python tools/glgm06_probe.py 2000 0 3 2>&1 | grep run_ms | tail -1 | cut -c1-230; fi
done
echo "== dense, chain not exclusive"; for K in 4 8 16; do echo "reserve $K"; REACH_CHAIN_EXCLUSIVE=0 REACH_GEMM_RESERVE=$K REACH_GLGM06_DENSE=1 timeout 300 python tools/glgm06_probe.py 2000 0 3 2>&1 | grep run_ms | tail -1 | cut -c1-230; done
echo "== sparse, chain not exclusive"; REACH_CHAIN_EXCLUSIVE=0 timeout 300 python tools/glgm06_probe.py 2000 0 3 2>&1 | grep run_ms | tail -1 | cut -c1-230
