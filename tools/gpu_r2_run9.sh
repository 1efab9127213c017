#!/bin/bash
# 8-GPU box: delivery check at world 8, bench at N=8 and N=4 (and N=2 with the new knobs)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_check.py > gpurun_out/r2_9_multi.log 2>&1; grep -E "MULTI_CHECK|False" gpurun_out/r2_9_multi.log | head
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_9_bench_n$N.json 2> gpurun_out/r2_9_bench_n$N.err; tail -2 gpurun_out/r2_9_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_9_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N", {k:round(d[k],2) for k in ("value","ms_per_step","fps","fps_at_2_in_flight","fps_latency")}, "e2e fps", round(d["e2e"]["fps"],1), "cfg5", d.get("config5"))
PY
done
