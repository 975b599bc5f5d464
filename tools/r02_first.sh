#!/bin/bash
# Round 2, first GPU call: where does the cost of smaller staging blocks come from; final-build cuBLAS A/B.
set -u
mkdir -p gpurun_out
for br in 8192 4096 2048; do
  python bench.py --no-salun --no-cpu-baseline --steps 10 --warmup 3 --block-rows $br > gpurun_out/r02_blk$br.json 2> gpurun_out/r02_blk$br.err
  tail -c 1500 gpurun_out/r02_blk$br.json
done
python tools/gpu_cublas_ab.py > gpurun_out/r02_cublas_ab.log 2>&1; tail -8 gpurun_out/r02_cublas_ab.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_blk2048.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-salun --block-rows 2048 > gpurun_out/r02_ncu2048.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_blk8192.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-salun > gpurun_out/r02_ncu8192.log 2>&1
ls gpurun_out | tail -5
