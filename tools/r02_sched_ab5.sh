#!/bin/bash
# step-level A/B, alternating processes on one box: the library before (old) and after (final) the scheduler-ring / second TMA thread work
set -u
O=gpurun_out
L=multimodal_unlearning_b200
F="--steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras"
for i in 1 2 3 4; do
  for v in old final; do
    cp $L/_ab/$v.so $L/libclear_b200.so
    python bench.py $F > $O/r02_sab5_${v}_$i.json 2> $O/r02_sab5_${v}_$i.err
    python - <<PY
import json
d=json.loads(open("$O/r02_sab5_${v}_$i.json").read().strip().splitlines()[-1])
print("$v", $i, round(d['ms_per_step'],3), round(d['value_with_stage_events']['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['frac'],3), {k: round(v[1],3) for k,v in d['roofline']['stages_ms_per_step'].items()}, d['clocks']['sm_mhz'])
PY
  done
done
cp $L/_ab/final.so $L/libclear_b200.so
