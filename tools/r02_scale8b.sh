#!/bin/bash
# final build on one 8-GPU box: 1 / 4 / 8 GPUs with the fused all-reduce, 8 GPUs with NCCL, 8 GPUs on BASELINE configs[4]
set -u
F="--steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533"
python bench.py --gpus 1 $F > gpurun_out/r02_s8b_n1.json 2> gpurun_out/r02_s8b_n1.err
for n in 4 8; do
  $TR --nproc-per-node $n bench.py --gpus $n $F > gpurun_out/r02_s8b_n$n.json 2> gpurun_out/r02_s8b_n$n.err; echo "n=$n rc=$?"
done
$TR --nproc-per-node 8 bench.py --gpus 8 $F --allreduce nccl > gpurun_out/r02_s8b_n8_nccl.json 2> gpurun_out/r02_s8b_n8_nccl.err; echo "nccl rc=$?"
$TR --nproc-per-node 8 bench.py --gpus 8 $F --config 5 > gpurun_out/r02_s8b_cfg5_n8.json 2> gpurun_out/r02_s8b_cfg5_n8.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
for f in ("n1","n4","n8","n8_nccl","cfg5_n8"):
    try:
        d=json.loads(open(f"gpurun_out/r02_s8b_{f}.json").read().strip().splitlines()[-1])
        r=d['roofline']
        print(f, round(d['value']), round(d['ms_per_step'],3), round(d['value_with_stage_events']['ms_per_step'],3), round(d['e2e']['value']), round(r['frac'],3), round(r['kernel_ms_per_step'],3), round(r['kernel_ms_per_step_fastest_rank'],3), d.get('dw_allreduce_check',{}).get('rel'), d.get('dw_allreduce_check',{}).get('identical_across_ranks'), d['clocks']['sm_mhz'])
    except Exception as e: print(f, "ERR", e)
PY
tail -c 300 gpurun_out/r02_s8b_n8.err
