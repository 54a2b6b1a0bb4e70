#!/bin/bash
# One gpurun call: GPU tests, default bench, then the C4 (n=10913) bench on the GPUs of this box ($1 = number of GPUs).
mkdir -p gpurun_out
G="${1:-1}"
if [ "$G" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -3 gpurun_out/pytest_gpu.log
  timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
  tail -2 gpurun_out/bench.log | cut -c1-600
  timeout 1500 python bench.py --workload C4 --steps 1 --warmup 3 > gpurun_out/bench_c4_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c4_1gpu.log
  tail -2 gpurun_out/bench_c4_1gpu.log
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $G --workload C4 --steps 1 --warmup 3 > gpurun_out/bench_c4_${G}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c4_${G}gpu.log
  tail -2 gpurun_out/bench_c4_${G}gpu.log
fi
