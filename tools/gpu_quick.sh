#!/bin/bash
# One gpurun call: GPU parity tests, then the default bench (C3 + C2 summary).  $1 = optional pytest -k expression.
mkdir -p gpurun_out
K="${1:-}"
timeout 1200 python -m pytest tests -m gpu -x -q ${K:+-k "$K"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
tail -2 gpurun_out/bench.log
