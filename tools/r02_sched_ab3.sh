#!/bin/bash
set -u
O=gpurun_out
L=multimodal_unlearning_b200
for v in "$@"; do
  cp $L/_ab/$v.so $L/libclear_b200.so
  ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 -s 5 -c 1 -o $O/r02_sab3_$v -f python tools/gpu_gemm_shapes.py 3 > /dev/null 2>&1
  ncu -i $O/r02_sab3_$v.ncu-rep --page source --csv --print-source sass > $O/r02_sab3_${v}_src.csv 2>/dev/null
  rm -f $O/r02_sab3_$v.ncu-rep
done
ls -la $O | grep sab3
