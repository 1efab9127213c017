#!/bin/bash
# under `gpurun --gpus 2`: the cost-balanced partition on two real GPUs (frames identical to single-GPU frames), then one-GPU emulation of
# every rank of an 8-GPU job with the partition $PART (interleaved | balanced | contiguous)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_check.py > gpurun_out/r2_28_multi_n2.log 2>&1; grep -E "MULTI_CHECK|False|Error" gpurun_out/r2_28_multi_n2.log | head -3
emu() {  # rank partition device
  CUDA_VISIBLE_DEVICES=$3 GRT_BENCH_FRAMES_IN_FLIGHT=12 GRT_BENCH_BLOCKS_PER_SM=1 GRT_BENCH_PARTITION=$2 GRT_BENCH_EMULATE_WORLD=8 GRT_BENCH_EMULATE_RANK=$1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side-measurements 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('world 8 rank $1 partition $2: ms_per_step %.2f e2e_fps %.1f rank steps %.0f M' % (d['ms_per_step'], d['e2e']['fps'], d['frame']['sum_ranks_ray_steps'] / 1e6), d['frame'].get('bounds'), (d['frame'].get('block_costs_mray_steps_by_round') or [None])[-1])
" | tee -a gpurun_out/r2_28_emulate_$3.log
}
(for r in 0 2 4 6; do emu $r ${PART:-contiguous} 0; done) &
(for r in 1 3 5 7; do emu $r ${PART:-contiguous} 1; done) &
wait
cat gpurun_out/r2_28_emulate_0.log gpurun_out/r2_28_emulate_1.log | sort
