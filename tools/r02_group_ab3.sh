#!/bin/bash
# compacted CE steps: dH + dW grouped (default) against separate launches (dH with the in-kernel split-K choice over an fp32 scratch)
F="--steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras"
for i in 1 2; do
for g in 2 0; do for c in 4 3; do
  CLEAR_GEMM_GROUP=$g python bench.py $F --config $c > gpurun_out/r02_gab3.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_gab3.json").read().strip().splitlines()[-1])
print("group=$g cfg$c", round(d['ms_per_step'],3), round(d['value_with_stage_events']['ms_per_step'],3), {k: round(v[1],3) for k,v in d['roofline']['stages_ms_per_step'].items()})
PY
done; done; done
