#!/bin/bash
# NPO row pass in isolation (N = 4096 rows): library builds interleaved across processes, then instruction counts under ncu
set -u
L=multimodal_unlearning_b200
for i in 1 2 3; do
  for v in "$@"; do
    cp $L/_ab/$v.so $L/libclear_b200.so
    python tools/gpu_rowloss_ab.py $v 2>&1 | grep -E " npo | kl_rev | ce " | tr '\n' ' '; echo
  done
done
for v in "$@"; do
  cp $L/_ab/$v.so $L/libclear_b200.so
  ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active -k regex:rowloss_npo1 -c 2 --clock-control none --csv --log-file gpurun_out/r02_row_$v.csv python tools/gpu_rowloss_ab.py $v > /dev/null 2>&1
  python - $v <<'PY'
import csv, sys
v=sys.argv[1]
rows=[r for r in csv.reader(open(f"gpurun_out/r02_row_{v}.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r))
    if d['ID']=='1': print(v, d['Metric Name'], d['Metric Value'])
PY
done
