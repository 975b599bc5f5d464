#!/bin/bash
set -u
python -m pytest tests/test_salun_gpu.py -q -m gpu -k "adamw8" 2>&1 | tail -15 > gpurun_out/r02_tests_adamw8.log; tail -5 gpurun_out/r02_tests_adamw8.log
python tools/gpu_block_ab.py > gpurun_out/r02_block_ab.log 2>&1; tail -3 gpurun_out/r02_block_ab.log
python tools/gpu_compact_profile.py && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_compact_launches.csv python tools/gpu_compact_profile.py > gpurun_out/r02_compact_ncu.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-salun --no-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|rowloss' -s 22 -c 6 -o gpurun_out/r02_prof $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
