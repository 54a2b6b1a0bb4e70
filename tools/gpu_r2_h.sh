#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "glgm06 or api or smoke" > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_h.log
tail -15 gpurun_out/pytest_h.log
timeout 300 python tools/glgm06_probe.py 2000 0 3 2>&1 | tail -5
REACH_CHAIN_SEQUENTIAL=1 timeout 300 python tools/glgm06_probe.py 2000 0 3 2>&1 | tail -2
