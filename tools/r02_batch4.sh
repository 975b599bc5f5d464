#!/bin/bash
set -u
python -m pytest tests/test_head_gpu.py tests/test_trainer_gpu.py -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r02_tests_grouped.log; tail -4 gpurun_out/r02_tests_grouped.log
python tools/nvlink_probe.py > gpurun_out/r02_nvlink_probe.log 2>&1
python tools/gpu_block_ab.py 8192,4096,2048 > gpurun_out/r02_block_ab_grouped.log 2>&1; tail -2 gpurun_out/r02_block_ab_grouped.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-salun > gpurun_out/r02_bench_grouped.json 2> gpurun_out/r02_bench_grouped.err; tail -c 300 gpurun_out/r02_bench_grouped.err
