#!/bin/bash
# Sweep the two scheduling knobs of bench.py (frame slots in flight x integrate blocks per SM and slot); one JSON line each.
# usage (under gpurun): bash tools/gpu_sweep_bench.sh <tag> "<fif:bps> <fif:bps> ..."
tag=${1:-sweep}
mkdir -p gpurun_out
: > gpurun_out/sweep_$tag.jsonl
for combo in $2; do
  fif=${combo%%:*}; bps=${combo##*:}
  GRT_BENCH_FRAMES_IN_FLIGHT=$fif GRT_BENCH_BLOCKS_PER_SM=$bps timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null >> gpurun_out/sweep_$tag.jsonl
  tail -1 gpurun_out/sweep_$tag.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$combo', 'ms/frame', round(d['ms_per_step'],2), 'fps', round(d['fps'],2), 'latency', round(d['frame_latency_ms'],2), 'e2e fps', round(d['e2e']['fps'],2), 'stages', d['frame']['rank0_stage_ms']['prepass'], d['frame']['rank0_stage_ms']['integrate'], d['frame']['rank0_stage_ms']['integrate_extra'], d['kernel_info'])"
done
