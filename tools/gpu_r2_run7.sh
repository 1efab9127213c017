#!/bin/bash
# one-GPU emulation of rank 0 of an 8-GPU (and 4-GPU) job: which slots:blocks setting gives the best ms/frame at K=20
mkdir -p gpurun_out
for W in 8 4 2; do
for S in 4 6 8; do for B in 1 2 3; do
  GRT_BENCH_EMULATE_WORLD=$W GRT_BENCH_FRAMES_IN_FLIGHT=$S GRT_BENCH_BLOCKS_PER_SM=$B timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side-measurements 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('world $W slots $S blocks $B: ms_per_step %.2f fps2 %.1f latency_ms %.2f stages %s' % (d['ms_per_step'], d['fps_at_2_in_flight'], d['frame_latency_ms'], {k:round(v,2) for k,v in d['frame']['rank0_stage_ms_one_frame_alone'].items() if v>0.1}))
" | tee -a gpurun_out/r2_7_emulate.log
done; done; done
