#!/bin/bash
set -u
python -m pytest tests/test_salun_gpu.py tests/test_trainer_gpu.py -q -m gpu 2>&1 | tail -4
python tools/bench_salun.py --reps 3 > gpurun_out/r02_salun_bench3.log 2>&1; grep -A3 "8bit" gpurun_out/r02_salun_bench3.log | head -5
