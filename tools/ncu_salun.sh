#!/bin/bash
# Run under gpurun: plain run first (must exit 0), then one --set full capture of the SalUn / update kernels.
set -u
CMD="python tools/bench_salun.py --layers 4 --reps 1"
$CMD > gpurun_out/salun_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'histogram_kernel|mask_kernel|adamw_kernel|sqnorm_kernel|l1l0_kernel|accumulate_kernel' -c 24 -o gpurun_out/prof_salun $CMD > gpurun_out/ncu_salun.log 2>&1
tail -3 gpurun_out/ncu_salun.log
