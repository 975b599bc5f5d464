#!/bin/bash
# Everything the round's profiles/ are made from, in one gpurun call (1 GPU).  Outputs under gpurun_out/.
set -u
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/bench_configs.py > gpurun_out/configs_bench.log 2>&1; tail -1 gpurun_out/configs_bench.log | cut -c1-200
python tools/bench_salun.py --reps 3 > gpurun_out/salun_bench.log 2>&1
bash tools/ncu_round.sh > gpurun_out/ncu_round.log 2>&1
bash tools/ncu_salun.sh > gpurun_out/ncu_salun_round.log 2>&1
ls gpurun_out | head -80
