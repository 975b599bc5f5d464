#!/bin/bash
# Run under gpurun: plain bench (must exit 0) -> launch list -> one full capture of the hot kernels.
set -u
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-salun"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|rowloss' -s 22 -c 6 -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
