#!/bin/bash
# session-2 helper: GLGM06 tests + C2 sparse / dense probes with a reserve sweep
python -m pytest tests/test_glgm06_gpu.py tests/test_glgm06_full_gpu.py -m gpu -x -q > gpurun_out/s2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2a_pytest.log
tail -2 gpurun_out/s2a_pytest.log
echo "== sparse"; python tools/glgm06_probe.py 2000 0 3 2>&1 | grep steps_per_s | tail -1
echo "== sparse one-stream"; REACH_NUMERICS_ONE_STREAM=1 python tools/glgm06_probe.py 2000 0 3 2>&1 | grep steps_per_s | tail -1
echo "== dense default"; REACH_GLGM06_DENSE=1 python tools/glgm06_probe.py 2000 0 3 2>&1 | grep steps_per_s | tail -1
echo "== dense one-stream"; REACH_NUMERICS_ONE_STREAM=1 REACH_GLGM06_DENSE=1 python tools/glgm06_probe.py 2000 0 3 2>&1 | grep steps_per_s | tail -1
for K in 6 8 10 12 16 20; do echo "== dense RESERVE $K"; REACH_GLGM06_DENSE=1 REACH_GEMM_RESERVE=$K python tools/glgm06_probe.py 2000 0 3 2>&1 | grep steps_per_s | tail -1; done
