#!/bin/bash
# large dense GLGM06 on 2 GPUs: batch size sweep for the pipelined column-sharded engine
for S in 4 8 16 32 64; do
  echo "== BLOCK $S"
  REACH_GLGM06_BLOCK=$S timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29520+S)) tools/glgm06_large_sharded.py 512 200 10 2>&1 | tail -1
done
