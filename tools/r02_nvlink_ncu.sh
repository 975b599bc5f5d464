#!/bin/bash
# gpurun --gpus 2: plain run first, then ncu with whatever NVLink byte counters this chip exposes
set -u
python tools/peer_dw_probe.py > gpurun_out/r02_peer_probe.log 2>&1 || { tail -5 gpurun_out/r02_peer_probe.log; exit 1; }
tail -4 gpurun_out/r02_peer_probe.log
ncu --query-metrics 2>/dev/null | grep -i "^nvl" > gpurun_out/r02_nvl_metrics.txt
wc -l gpurun_out/r02_nvl_metrics.txt
M=$(awk '{print $1}' gpurun_out/r02_nvl_metrics.txt | grep -E "bytes" | grep -vE "pct|per_second" | head -24 | sed 's/$/.sum/' | paste -sd, -)
echo "metrics: $M"
ncu --metrics gpu__time_duration.sum,$M --clock-control none -k regex:gemm_bf16 -c 4 --csv --log-file gpurun_out/r02_nvlink_ncu.csv \
  python tools/peer_dw_probe.py > gpurun_out/r02_nvlink_ncu.log 2>&1
tail -3 gpurun_out/r02_nvlink_ncu.log; head -c 1500 gpurun_out/r02_nvlink_ncu.csv
