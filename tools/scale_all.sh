#!/bin/bash
# 1, 2, 4, 8 GPUs back to back on one 8-GPU box (the driver's scaling run, reproduced): fused all-reduce (default).
python bench.py --gpus 1 --steps 20 --warmup 5 --no-salun --no-cpu-baseline 2> gpurun_out/sa_1.err | tail -1 > gpurun_out/sa_1.json
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
    bench.py --gpus $n --steps 20 --warmup 5 --no-salun --no-cpu-baseline 2> gpurun_out/sa_$n.err | tail -1 > gpurun_out/sa_$n.json
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/sa_{n}.json").read())
        base = base or d["value"]
        print(f"| {n} | {d['value']:.0f} | {d['ms_per_step']:.3f} | {d['e2e']['value']:.0f} | {d['value']/base:.3f} | {d['value']/base/n:.3f} | {d['clocks']['sm_mhz']} |")
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/sa_{n}.err").read()[-1500:])
PY
