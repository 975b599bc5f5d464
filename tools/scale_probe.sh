#!/bin/bash
# Diagnostic: where does the multi-GPU step lose time?  argv[1] = number of GPUs.
N=${1:-2}
run() {
  tag=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-salun --no-cpu-baseline $EXTRA > gpurun_out/scale_$tag.json 2> gpurun_out/scale_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "ms/step", round(d["ms_per_step"], 3), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gemm ms", round(d["roofline"]["kernel_ms_per_step"], 3), "rowloss", round(d["roofline"]["rowloss_ms_per_step"], 3), d["clocks"])
except Exception as e:
    print("$tag failed", e)
    print(open("gpurun_out/scale_$tag.err").read()[-1500:])
PY
}
EXTRA="--allreduce fused" run fused A=1
EXTRA="--no-allreduce" run noallreduce A=1
EXTRA="--allreduce nccl" run nccl A=1
EXTRA="--allreduce fused" run fused2 A=1
