#!/bin/bash
# 8-GPU box: delivery check, C++ host over 8 devices, bench at N=8, 4, 2, 1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_check.py > gpurun_out/r2_16_multi.log 2>&1; grep -E "MULTI_CHECK|False" gpurun_out/r2_16_multi.log | head -3
(cd /tmp && timeout 300 $GRAFT_REPO_ROOT/geodesic_raytracing_b200/grt_headless -content $GRAFT_REPO_ROOT/geodesic_raytracing_b200/content -bench kerr_schild -size 3840 2160 -frames 16 -inflight 8 -gpus 8 -out /tmp/f8.png > $GRAFT_REPO_ROOT/gpurun_out/r2_16_headless8.log 2>&1; timeout 300 $GRAFT_REPO_ROOT/geodesic_raytracing_b200/grt_headless -content $GRAFT_REPO_ROOT/geodesic_raytracing_b200/content -bench kerr_schild -size 3840 2160 -frames 8 -inflight 4 -gpus 1 -out /tmp/f1.png > $GRAFT_REPO_ROOT/gpurun_out/r2_16_headless1.log 2>&1; cmp /tmp/f8.png /tmp/f1.png && echo "headless: 8-GPU frame == 1-GPU frame")
tail -3 gpurun_out/r2_16_headless8.log; tail -2 gpurun_out/r2_16_headless1.log
for N in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_16_bench_n$N.json 2> gpurun_out/r2_16_bench_n$N.err; tail -2 gpurun_out/r2_16_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_16_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N", {k:round(d[k],2) for k in ("value","ms_per_step","fps","fps_at_2_in_flight","fps_latency")}, "e2e fps", round(d["e2e"]["fps"],1), "by rank", [round(x,1) for x in d["frame"]["timed_region_ms_by_rank"]], "cfg5 fps", d.get("config5",{}).get("fps"))
PY
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_16_bench_n1.json 2> gpurun_out/r2_16_bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_16_bench_n1.json").read().strip().splitlines()[-1])
print("N=1", {k:round(d[k],2) for k in ("value","ms_per_step","fps","fps_at_2_in_flight","fps_latency")}, "e2e fps", round(d["e2e"]["fps"],1), "cfg5 fps", d.get("config5",{}).get("fps"))
PY
