#!/bin/bash
# Round 2: the driver's bench command, then the staging-block variants through the one-call entry (value_one_call).
set -u
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "rc=$?"; tail -c 600 gpurun_out/r02_bench_n1.err
for br in 8192 4096 2048 1024; do
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras --block-rows $br > gpurun_out/r02_blk${br}.json 2> gpurun_out/r02_blk${br}.err
done
for c in 3 4 5; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-salun --no-extras --config $c > gpurun_out/r02_cfg${c}.json 2> gpurun_out/r02_cfg${c}.err
done
python bench.py --impl reference --steps 3 --warmup 1 --config 5 > gpurun_out/r02_ref_cfg5.json 2>&1
ls -la gpurun_out | tail -12
