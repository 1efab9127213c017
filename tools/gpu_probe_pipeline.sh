#!/bin/bash
# Diagnostic: the per-rank pipeline of an N-GPU job emulated on ONE GPU (GRT_BENCH_EMULATE_WORLD: rank 0's stripes, no gather),
# for several (frames in flight : blocks per SM and slot : timed frames).
# usage (under gpurun): bash tools/gpu_probe_pipeline.sh <emulated world> "<fif:bps:K> ..."
w=$1
for combo in $2; do
  IFS=: read fif bps k <<< "$combo"
  GRT_BENCH_EMULATE_WORLD=$w GRT_BENCH_FRAMES_IN_FLIGHT=$fif GRT_BENCH_BLOCKS_PER_SM=$bps timeout 300 python bench.py --steps $k --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('emulated world $w  fif $fif bps $bps K $k', 'ms/frame', round(d['ms_per_step'],2), 'latency', round(d['frame_latency_ms'],2), 'e2e fps', round(d['e2e']['fps'],2), d['frame']['rank0_stage_ms']['prepass'], d['frame']['rank0_stage_ms']['integrate'], d['frame']['rank0_stage_ms']['integrate_extra'])"
done
