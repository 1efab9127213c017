#!/bin/bash
# One-GPU emulation of rank R of an N-GPU job (no delivery): does the number of hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS, default 8)
# limit the overlap of the frame slots' streams?  usage (under gpurun): bash tools/gpu_r2_emulate_connections.sh ['world rank slots blocks connections' lines]
mkdir -p gpurun_out
run() {  # world rank slots blocks connections
  CUDA_DEVICE_MAX_CONNECTIONS=$5 GRT_BENCH_EMULATE_WORLD=$1 GRT_BENCH_EMULATE_RANK=$2 GRT_BENCH_FRAMES_IN_FLIGHT=$3 GRT_BENCH_BLOCKS_PER_SM=$4 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side-measurements 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('world $1 rank $2 slots $3 blocks $4 connections $5: ms_per_step %.2f latency %.2f e2e_fps %.1f' % (d['ms_per_step'], d['frame_latency_ms'], d['e2e']['fps']))
" | tee -a gpurun_out/r2_26_emulate.log
}
if [ -n "$1" ]; then
  while read -r w r sl b c; do [ -n "$w" ] && run $w $r $sl $b $c; done <<< "$1"
  exit 0
fi
run 8 0 8 1 8
run 8 0 8 1 32
run 8 0 12 1 32
run 8 0 16 1 32
run 8 3 8 1 8
run 8 3 12 1 32
run 1 0 4 3 8
run 1 0 4 3 32
run 1 0 6 3 32
run 1 0 8 2 32
