#!/bin/bash
# old (fenced scheduler ring, every reader decodes) vs new (st.async ring with decoded items, fence-free hand-backs) builds of
# the library, alternating processes on one box; then the GPU tests on the new build and the tensor-pipe counter of its projection
set -u
O=gpurun_out
L=multimodal_unlearning_b200
F="--steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras"
cp $L/_ab/new.so $L/libclear_b200.so
python -m pytest tests/test_head_gpu.py -m gpu -q -x 2>&1 | tail -2
for i in 1 2 3; do
  for v in old new; do
    cp $L/_ab/$v.so $L/libclear_b200.so
    python bench.py $F > $O/r02_sab_${v}_$i.json 2> $O/r02_sab_${v}_$i.err
    python bench.py $F --config 5 > $O/r02_sab5_${v}_$i.json 2>> $O/r02_sab_${v}_$i.err
    python - <<PY
import json
for f in ("$O/r02_sab_${v}_$i.json", "$O/r02_sab5_${v}_$i.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print("$v", $i, d['config']['baseline_config'], round(d['ms_per_step'],3), round(d['value_with_stage_events']['ms_per_step'],3), {k: round(v[1],3) for k,v in d['roofline']['stages_ms_per_step'].items()}, d['clocks']['sm_mhz'])
PY
  done
done
cp $L/_ab/new.so $L/libclear_b200.so
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_bf16 -s 8 -c 8 --csv --log-file $O/r02_sab_ncu_new.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-salun --no-extras > /dev/null 2>&1
cp $L/_ab/old.so $L/libclear_b200.so
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_bf16 -s 8 -c 8 --csv --log-file $O/r02_sab_ncu_old.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-salun --no-extras > /dev/null 2>&1
cp $L/_ab/new.so $L/libclear_b200.so
python - <<'PY'
import csv
for v in ("old","new"):
    rows=[r for r in csv.reader(open(f"gpurun_out/r02_sab_ncu_{v}.csv")) if len(r)>10]
    h=rows[0]
    for r in rows[1:]:
        d=dict(zip(h,r)); print(v, d['Kernel Name'][:45], d['Metric Name'][:40], d['Metric Value'])
PY
