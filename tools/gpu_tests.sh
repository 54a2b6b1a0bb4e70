#!/bin/bash
# One gpurun call: GPU parity tests (optionally a subset: $1 = pytest -k expression), then smoke.
mkdir -p gpurun_out
K="${1:-}"
if [ -n "$K" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/pytest_gpu.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
fi
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
