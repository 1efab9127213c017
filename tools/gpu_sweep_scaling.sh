#!/bin/bash
# bench.py on N GPUs of one box for a list of scheduling settings (frame slots in flight : integrate blocks per SM and slot).
# usage (under gpurun --gpus N): bash tools/gpu_sweep_scaling.sh <N> "<fif:bps> ..."
n=$1
mkdir -p gpurun_out
: > gpurun_out/scale_sweep_n$n.jsonl
for combo in $2; do
  fif=${combo%%:*}; bps=${combo##*:}
  if [ "$n" = "1" ]; then
    GRT_BENCH_FRAMES_IN_FLIGHT=$fif GRT_BENCH_BLOCKS_PER_SM=$bps timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null >> gpurun_out/scale_sweep_n$n.jsonl
  else
    GRT_BENCH_FRAMES_IN_FLIGHT=$fif GRT_BENCH_BLOCKS_PER_SM=$bps timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 10 --warmup 3 2>/dev/null | grep '^{' >> gpurun_out/scale_sweep_n$n.jsonl
  fi
  tail -1 gpurun_out/scale_sweep_n$n.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$n $combo', 'ms/frame', round(d['ms_per_step'],2), 'fps', round(d['fps'],2), 'latency', round(d['frame_latency_ms'],2), 'e2e fps', round(d['e2e']['fps'],2))"
done
