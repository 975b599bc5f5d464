#!/bin/bash
# 2 GPUs: the data-parallel tests, then bench.py as the driver launches it (fused all-reduce), the reference arm under torchrun
set -u
O=gpurun_out
python -m pytest tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -3
F="--steps 20 --warmup 5"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
$TR --nproc-per-node 2 bench.py --gpus 2 $F > $O/r02_v2_n2.json 2> $O/r02_v2_n2.err; echo rc=$?
$TR --nproc-per-node 2 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > $O/r02_v2_ref.json 2> $O/r02_v2_ref.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_v2_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['value_with_stage_events']['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('dw_allreduce_check'), d['clocks'])
r=json.loads(open('gpurun_out/r02_v2_ref.json').read().strip().splitlines()[-1]); print(r['impl'], r['value'], r['n_gpus'])
PY
tail -c 300 $O/r02_v2_n2.err
