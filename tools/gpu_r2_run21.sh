#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_frame_probe.py --tag ballot2 --slots 1 2 4 --blocks 3 2>&1 | tee gpurun_out/r2_21_probe.log
timeout 900 python -m pytest tests -m gpu -q -x --durations=6 > gpurun_out/r2_21_pytest.log 2>&1; echo rc $? >> gpurun_out/r2_21_pytest.log; grep -E "passed|failed|FAILED|^rc|^E  |s call" gpurun_out/r2_21_pytest.log | head -20
for W in 8 4; do
  GRT_BENCH_EMULATE_WORLD=$W GRT_BENCH_FRAMES_IN_FLIGHT=8 GRT_BENCH_BLOCKS_PER_SM=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side-measurements 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('emulated rank 0 of $W: ms_per_step %.2f alone %.2f fps2 %.1f stages %s' % (d['ms_per_step'], d['frame']['rank0_frame_alone_ms'], d['fps_at_2_in_flight'], {k:round(v,2) for k,v in d['frame']['rank0_stage_ms_one_frame_alone'].items() if v>0.1}))
"
done
