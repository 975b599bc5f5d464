#!/bin/bash
# one 8-GPU box: 1 / 2 / 4 / 8 GPUs with the fused all-reduce, 8 GPUs with NCCL, 8 GPUs on BASELINE configs[4]
set -u
F="--steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533"
python bench.py --gpus 1 $F > gpurun_out/r02_s8_n1.json 2> gpurun_out/r02_s8_n1.err
for n in 2 4 8; do
  $TR --nproc-per-node $n bench.py --gpus $n $F > gpurun_out/r02_s8_n$n.json 2> gpurun_out/r02_s8_n$n.err; echo "n=$n rc=$?"
done
$TR --nproc-per-node 8 bench.py --gpus 8 $F --allreduce nccl > gpurun_out/r02_s8_n8_nccl.json 2> gpurun_out/r02_s8_n8_nccl.err; echo "nccl rc=$?"
$TR --nproc-per-node 8 bench.py --gpus 8 $F --no-allreduce > gpurun_out/r02_s8_n8_none.json 2> gpurun_out/r02_s8_n8_none.err; echo "none rc=$?"
python bench.py --gpus 1 $F --config 5 > gpurun_out/r02_s8_cfg5_n1.json 2> gpurun_out/r02_s8_cfg5_n1.err
$TR --nproc-per-node 8 bench.py --gpus 8 $F --config 5 > gpurun_out/r02_s8_cfg5_n8.json 2> gpurun_out/r02_s8_cfg5_n8.err; echo "cfg5 rc=$?"
tail -c 300 gpurun_out/r02_s8_n8.err
