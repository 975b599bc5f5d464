#!/bin/bash
# source-level stall samples of the final contraction kernel (projection pair and dH + dW... the three single launches of gpu_gemm_shapes.py)
set -u
O=gpurun_out
for idx in 3 5; do
  ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 -s $idx -c 1 -o $O/r02_fsrc_$idx -f python tools/gpu_gemm_shapes.py 3 > /dev/null 2>&1
  ncu -i $O/r02_fsrc_$idx.ncu-rep --page source --csv --print-source sass > $O/r02_fsrc_${idx}_src.csv 2>/dev/null
  rm -f $O/r02_fsrc_$idx.ncu-rep
done
ls -la $O | grep fsrc
