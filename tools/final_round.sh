#!/bin/bash
# Everything the round's profiles/ are made from, in one gpurun call (1 GPU).  Outputs under gpurun_out/ (prefix r02f_).
set -u
O=gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
python -m pytest tests -m gpu -q > $O/r02f_pytest_gpu.log 2>&1; tail -2 $O/r02f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02f_smoke.log 2>&1; tail -1 $O/r02f_smoke.log
fi
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r02f_bench_ref.json 2> $O/r02f_bench_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02f_bench.json 2> $O/r02f_bench.err; tail -c 300 $O/r02f_bench.json; echo
for c in 3 4 5; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-salun --no-extras --config $c > $O/r02f_cfg$c.json 2> $O/r02f_cfg$c.err; done
python tools/bench_configs.py > $O/r02f_configs.log 2>&1; cp $O/configs_bench.md $O/r02f_configs.md 2>/dev/null
python tools/gpu_block_ab.py 8192,4096,2048,1024 > $O/r02f_block_ab.log 2>&1
# ncu: launch list of the bench command, then one full capture of the hot kernels (each only after the plain run exited 0)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-salun --no-extras"
$CMD > $O/r02f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f_bench_launches.csv $CMD > $O/r02f_ncu_list.log 2>&1
$CMD > $O/r02f_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|rowloss' -s 12 -c 6 -o $O/r02f_prof $CMD > $O/r02f_ncu_full.log 2>&1
ncu -i $O/r02f_prof.ncu-rep --page raw --csv > $O/r02f_prof_raw.csv 2>/dev/null
SAL="python tools/bench_salun.py --layers 4 --reps 1"
$SAL > $O/r02f_salun_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'histogram_kernel|hist_total|select_global|mask_kernel|adamw_kernel|adamw8_kernel|sqnorm_kernel|l1l0_kernel|accumulate_kernel' -c 30 -o /tmp/r02f_prof_salun $SAL > $O/r02f_ncu_salun.log 2>&1
ncu -i /tmp/r02f_prof_salun.ncu-rep --page raw --csv > $O/r02f_prof_salun_raw.csv 2>/dev/null   # the report itself stays on the box (64 MiB limit)
du -sh $O
ls -la $O | grep r02f_ | awk '{print $5, $9}'
