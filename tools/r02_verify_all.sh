#!/bin/bash
# everything the driver runs at round end, on a 2-GPU box: all GPU tests (incl. the 2-rank ones), smoke, bench at N = 1 and N = 2
set -u
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
F="--steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras"
python bench.py --gpus 1 $F > $O/r02_va_n1.json 2> $O/r02_va_n1.err; echo rc=$?
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
$TR --nproc-per-node 2 bench.py --gpus 2 $F > $O/r02_va_n2.json 2> $O/r02_va_n2.err; echo rc=$?
$TR --nproc-per-node 2 bench.py --gpus 2 $F --config 5 > $O/r02_va_n2c5.json 2> $O/r02_va_n2c5.err; echo rc=$?
python - <<'PY'
import json
for f in ("r02_va_n1","r02_va_n2","r02_va_n2c5"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],3), round(d['value_with_stage_events']['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), round(d['roofline']['frac'],3), {k: round(v[1],3) for k,v in d['roofline']['stages_ms_per_step'].items()}, d.get('dw_allreduce_check',{}).get('rel'), d['clocks']['sm_mhz'])
PY
