#!/bin/bash
# round 2, GPU call 5 (2 GPUs): GPU tests, parity report, new bench at N=1, multi-process delivery check and bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_5_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_5_pytest.log
grep -E "passed|failed|FAILED|rc" gpurun_out/r2_5_pytest.log
timeout 300 python tools/gpu_parity_report.py 256 144 > gpurun_out/r2_5_parity.log 2>&1; cat gpurun_out/r2_5_parity.log | cut -c1-150
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_5_bench_n1.json 2> gpurun_out/r2_5_bench_n1.err; tail -3 gpurun_out/r2_5_bench_n1.err; cut -c1-1200 gpurun_out/r2_5_bench_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_check.py > gpurun_out/r2_5_multi.log 2>&1; tail -15 gpurun_out/r2_5_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_5_bench_n2.json 2> gpurun_out/r2_5_bench_n2.err; tail -3 gpurun_out/r2_5_bench_n2.err; cut -c1-700 gpurun_out/r2_5_bench_n2.json
