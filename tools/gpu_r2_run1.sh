#!/bin/bash
# round 2, GPU call 1: tests, baseline bench, occupancy-tier sweep (lone-warp step latency shows in the prepass time), config 5 probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/r2_1_smi.txt
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_1_pytest.log
tail -5 gpurun_out/r2_1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err; tail -c 600 gpurun_out/r2_1_bench.json
for MB in 1 3 5; do
  GRT_INTEGRATE_MIN_BLOCKS=$MB GRT_CACHE_DIR=/tmp/cache_mb$MB timeout 300 python tools/gpu_frame_probe.py --tag mb$MB --slots 1 4 --blocks 1 3 >> gpurun_out/r2_1_tiers.log 2>&1
done
timeout 300 python tools/gpu_frame_probe.py --tag tier7 --slots 1 2 4 --blocks 1 3 0 >> gpurun_out/r2_1_tiers.log 2>&1
timeout 600 python tools/gpu_frame_probe.py --tag cfg5 --metric double_unequal_kerr --slots 1 2 4 --blocks 0 3 2 1 --frames 8 >> gpurun_out/r2_1_cfg5.log 2>&1
cat gpurun_out/r2_1_tiers.log gpurun_out/r2_1_cfg5.log
