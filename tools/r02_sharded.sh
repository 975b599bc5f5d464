#!/bin/bash
set -u
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533 --nproc-per-node $N tools/bench_sharded_masks.py > gpurun_out/r02_sharded_n$N.json 2> gpurun_out/r02_sharded_n$N.err; echo rc=$?
tail -c 600 gpurun_out/r02_sharded_n$N.err; cat gpurun_out/r02_sharded_n$N.json
