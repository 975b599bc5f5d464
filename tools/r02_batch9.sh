#!/bin/bash
# bench.py after the headline loop lost its stage events: per-step times of both loops, configs 5 (twice) and 2
O=gpurun_out
for c in 5 5 2 4; do
CLEAR_BENCH_DEBUG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-salun --no-extras --config $c > $O/r02_b9_cfg$c.json 2> $O/r02_b9_cfg$c.err
tail -2 $O/r02_b9_cfg$c.err | cut -c1-400
python -c "
import json; d=json.loads(open('$O/r02_b9_cfg$c.json').read().strip().splitlines()[-1])
print($c, d['ms_per_step'], d['value_with_stage_events']['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['clocks'])"
done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-salun --no-extras > $O/r02_b9_head.json 2> $O/r02_b9_head.err
python -c "
import json; d=json.loads(open('$O/r02_b9_head.json').read().strip().splitlines()[-1])
print('head', d['ms_per_step'], d['value_with_stage_events']['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['clocks'])"
