#!/bin/bash
# same box: 1 GPU, then 2 GPUs with the fused all-reduce, then 2 GPUs with NCCL
set -u
F="--steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras"
python bench.py --gpus 1 $F > gpurun_out/r02_s_n1.json 2> gpurun_out/r02_s_n1.err
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
$TR --nproc-per-node 2 bench.py --gpus 2 $F > gpurun_out/r02_s_n2.json 2> gpurun_out/r02_s_n2.err; echo rc=$?
$TR --nproc-per-node 2 bench.py --gpus 2 $F --allreduce nccl > gpurun_out/r02_s_n2_nccl.json 2> gpurun_out/r02_s_n2_nccl.err; echo rc=$?
$TR --nproc-per-node 2 bench.py --gpus 2 $F --no-allreduce > gpurun_out/r02_s_n2_none.json 2> gpurun_out/r02_s_n2_none.err; echo rc=$?
$TR --nproc-per-node 2 bench.py --gpus 2 $F --config 5 > gpurun_out/r02_s_n2_cfg5.json 2> gpurun_out/r02_s_n2_cfg5.err; echo rc=$?
tail -c 400 gpurun_out/r02_s_n2.err
