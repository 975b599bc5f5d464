#!/bin/bash
# 8 GPUs, same box: step time with the fused all-reduce, with no all-reduce at all (diagnostic), and with NCCL.
for mode in "--allreduce fused" "--no-allreduce" "--allreduce nccl"; do
  tag=$(echo $mode | tr -d ' -')
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --steps 20 --warmup 5 --no-salun --no-cpu-baseline $mode 2> gpurun_out/s8c_$tag.err | tail -1 > gpurun_out/s8c_$tag.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/s8c_$tag.json").read())
r = d["roofline"]
print("$tag", "ms/step", round(d["ms_per_step"], 3), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm max/min rank", round(r["kernel_ms_per_step"], 3), round(r["kernel_ms_per_step_fastest_rank"], 3))
PY
done
