#!/bin/bash
# grouped launches (student + teacher projection, dH + dW) against one launch per contraction, alternating processes on one box
for i in 1 2 3; do
  CLEAR_GEMM_GROUP=1 python tools/gpu_block_ab.py 8192,4096 2>&1 | sed 's/^/grouped  /'
  CLEAR_GEMM_GROUP=0 python tools/gpu_block_ab.py 8192,4096 2>&1 | sed 's/^/separate /'
done
