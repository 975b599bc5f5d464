#!/bin/bash
F="--steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras --config 5"
for i in 1 2 3 4 5 6 7 8; do
  CLEAR_BENCH_DEBUG=1 python bench.py $F 2>&1 >/dev/null | grep "headline" | cut -c1-420
done
