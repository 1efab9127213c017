#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list and one full capture of the integration kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_$tag.txt
nproc >> gpurun_out/gpu_$tag.txt
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu_$tag.log
  tail -5 gpurun_out/pytest_gpu_$tag.log
  timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_$tag.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_$tag.err | tee gpurun_out/bench_$tag.json | cut -c1-1500
tail -3 gpurun_out/bench_$tag.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2> gpurun_out/bench_ref_$tag.err | tee gpurun_out/bench_ref_$tag.json | cut -c1-900
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 3 -c 3 -o gpurun_out/prof_integrate_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out | tail -12
