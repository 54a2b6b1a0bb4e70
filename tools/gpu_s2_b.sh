#!/bin/bash
for K in 16 24 32 40 48; do echo "== dense EXCLUSIVE RESERVE $K"; REACH_GEMM_EXCLUSIVE=1 REACH_GLGM06_DENSE=1 REACH_GEMM_RESERVE=$K python tools/glgm06_probe.py 2000 0 3 2>&1 | grep steps_per_s | tail -1; done
