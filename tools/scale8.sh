#!/bin/bash
# 1 GPU and 8 GPUs back to back on the same box: fused vs NCCL all-reduce.
python bench.py --gpus 1 --steps 20 --warmup 5 --no-salun --no-cpu-baseline > gpurun_out/s8_n1.json 2> gpurun_out/s8_n1.err
for mode in fused nccl; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-salun --no-cpu-baseline --allreduce $mode > gpurun_out/s8_$mode.json 2> gpurun_out/s8_$mode.err
done
python - <<'PY'
import json
base = None
for tag in ("n1", "fused", "nccl"):
    try:
        d = json.loads(open(f"gpurun_out/s8_{tag}.json").read().strip().splitlines()[-1])
        if tag == "n1": base = d["value"]
        print(tag, "ms/step", round(d["ms_per_step"], 3), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "scaling x", round(d["value"] / base, 3) if base else None,
              "gemm", round(d["roofline"]["kernel_ms_per_step"], 3), d["clocks"])
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/s8_{tag}.err").read()[-2000:])
PY
