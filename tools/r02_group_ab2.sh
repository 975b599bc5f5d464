#!/bin/bash
# dH + dW as one launch at full waves (8192 rows) against two launches, with the fence-free scheduler ring; alternating processes
set -u
F="--steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras"
for i in 1 2 3 4; do
  for g in 1 2; do
    for c in 2 5; do
    CLEAR_GEMM_GROUP=$g python bench.py $F --config $c > gpurun_out/r02_gab_${g}_$i.json 2> gpurun_out/r02_gab_${g}_$i.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r02_gab_${g}_$i.json").read().strip().splitlines()[-1])
print("group=$g", $i, "cfg$c", round(d['ms_per_step'],3), round(d['value_with_stage_events']['ms_per_step'],3), {k: round(v[1],3) for k,v in d['roofline']['stages_ms_per_step'].items()}, d['clocks']['sm_mhz'])
PY
    done
  done
done
