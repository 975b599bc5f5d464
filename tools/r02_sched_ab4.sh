#!/bin/bash
set -u
L=multimodal_unlearning_b200
cp $L/_ab/two.so $L/libclear_b200.so
python -m pytest tests/test_head_gpu.py -m gpu -q -x 2>&1 | tail -2
bash tools/r02_sched_ab2.sh old new two two8
cp $L/_ab/two.so $L/libclear_b200.so
