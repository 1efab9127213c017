#!/bin/bash
mkdir -p gpurun_out
for ROWS in 54 36 108; do
for R in 0 1 2 3 4 5 6 7; do
  GRT_BENCH_STRIPE_ROWS=$ROWS GRT_BENCH_EMULATE_WORLD=8 GRT_BENCH_EMULATE_RANK=$R GRT_BENCH_FRAMES_IN_FLIGHT=8 GRT_BENCH_BLOCKS_PER_SM=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side-measurements 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rows $ROWS rank $R of 8: ms_per_step %.2f alone %.2f steps %d' % (d['ms_per_step'], d['frame']['rank0_frame_alone_ms'], d['frame']['sum_ranks_ray_steps']))
" | tee -a gpurun_out/r2_17_ranks.log
done; done
