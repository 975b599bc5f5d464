#!/bin/bash
set -u
python -m pytest tests/test_salun_gpu.py -q -m gpu -k "adamw8" 2>&1 | tail -4
python tools/bench_salun.py --reps 3 > gpurun_out/r02_salun_bench2.log 2>&1; grep -A6 "8bit" gpurun_out/r02_salun_bench2.log | head -8
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-salun --no-extras --config 5 > gpurun_out/r02_cfg5b.json 2> gpurun_out/r02_cfg5b.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_cfg5b.json').read().strip().splitlines()[-1]);print('cfg5',d['ms_per_step'],d['value_one_call']['ms_per_step'],d['roofline']['stages_ms_per_step'])"
SAL="python tools/bench_salun.py --layers 4 --reps 1"
ncu --set full --clock-control none -k regex:'adamw8_kernel' -c 2 -o /tmp/p8 $SAL > gpurun_out/r02_ncu_adamw8.log 2>&1
ncu -i /tmp/p8.ncu-rep --page raw --csv > gpurun_out/r02_adamw8_raw.csv 2>/dev/null; ls -la gpurun_out/r02_adamw8_raw.csv
