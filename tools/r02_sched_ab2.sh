#!/bin/bash
set -u
O=gpurun_out
L=multimodal_unlearning_b200
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,lts__t_sector_hit_rate.pct
for v in "$@"; do
  cp $L/_ab/$v.so $L/libclear_b200.so
  python tools/gpu_gemm_shapes.py 6 | sed "s/^/$v plain /"
  ncu --metrics $M --clock-control none -k regex:gemm_bf16 -s 3 -c 6 --csv --log-file $O/r02_sab2_$v.csv python tools/gpu_gemm_shapes.py 3 > /dev/null 2>&1
done
for v in "$@"; do
  cp $L/_ab/$v.so $L/libclear_b200.so
  python tools/gpu_gemm_shapes.py 6 | sed "s/^/$v plain2 /"
done
python - "$@" <<'PY'
import csv, sys, collections
for v in sys.argv[1:]:
    rows=[r for r in csv.reader(open(f"gpurun_out/r02_sab2_{v}.csv")) if len(r)>10]
    h=rows[0]; agg=collections.OrderedDict()
    for r in rows[1:]:
        d=dict(zip(h,r)); k=(d['Kernel Name'][22:34], d['Metric Name'][:28])
        agg.setdefault(k,[]).append(float(d['Metric Value'].replace(',','')))
    for k,vals in agg.items(): print(v, k[0], k[1], round(sum(vals)/len(vals),2))
PY
