#!/bin/bash
set -u
L=multimodal_unlearning_b200
for i in 1 2 3; do
  for v in "$@"; do
    cp $L/_ab/$v.so $L/libclear_b200.so
    python tools/gpu_rowloss_ab.py $v 2>&1 | grep -E " npo | kl_rev | ce | kl_fwd | entropy " | tr '\n' ' '; echo
  done
done
cp $L/_ab/row4.so $L/libclear_b200.so
python -m pytest tests/test_head_gpu.py -m gpu -q -x 2>&1 | tail -2
