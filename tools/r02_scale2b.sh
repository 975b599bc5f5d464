#!/bin/bash
# final build, one 2-GPU box: 1 GPU, 2 GPUs fused / NCCL / no all-reduce (diagnostic), alternating twice
set -u
F="--steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
for i in 1 2; do
python bench.py --gpus 1 $F > gpurun_out/r02_s2b_n1_$i.json 2>/dev/null
$TR --nproc-per-node 2 bench.py --gpus 2 $F > gpurun_out/r02_s2b_fused_$i.json 2>/dev/null
$TR --nproc-per-node 2 bench.py --gpus 2 $F --no-allreduce > gpurun_out/r02_s2b_none_$i.json 2>/dev/null
$TR --nproc-per-node 2 bench.py --gpus 2 $F --allreduce nccl > gpurun_out/r02_s2b_nccl_$i.json 2>/dev/null
done
python - <<'PY'
import json
for i in (1,2):
    for f in ("n1","fused","none","nccl"):
        d=json.loads(open(f"gpurun_out/r02_s2b_{f}_{i}.json").read().strip().splitlines()[-1]); r=d['roofline']
        print(i, f, round(d['value']), round(d['ms_per_step'],3), round(d['value_with_stage_events']['ms_per_step'],3), round(r['kernel_ms_per_step'],3), round(r['kernel_ms_per_step_fastest_rank'],3))
PY
