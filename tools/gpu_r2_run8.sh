#!/bin/bash
mkdir -p gpurun_out
for W in 8 4; do
for S in 8 12 16; do for B in 1 2; do
  GRT_BENCH_EMULATE_WORLD=$W GRT_BENCH_FRAMES_IN_FLIGHT=$S GRT_BENCH_BLOCKS_PER_SM=$B timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side-measurements 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('world $W slots $S blocks $B: ms_per_step %.2f e2e_fps %.1f' % (d['ms_per_step'], d['e2e']['fps']))
" | tee -a gpurun_out/r2_8_emulate.log
done; done; done
for S in 4 6 8; do
  GRT_BENCH_FRAMES_IN_FLIGHT=$S GRT_BENCH_BLOCKS_PER_SM=3 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side-measurements 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=1 slots $S blocks 3: ms_per_step %.2f e2e_fps %.1f' % (d['ms_per_step'], d['e2e']['fps']))
" | tee -a gpurun_out/r2_8_emulate.log
done
