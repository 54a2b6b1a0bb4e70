#!/bin/bash
# large dense GLGM06 (n = 1024): pipelined column-sharded engine at 2 / 4 / 8 GPUs
for W in 2 4 8; do
  echo "== WORLD $W"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29530+W)) tools/glgm06_large_sharded.py 512 200 10 2>&1 | tail -1
done
