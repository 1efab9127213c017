#!/bin/bash
# Multi-GPU visit (under `gpurun --gpus N`): delivery check at world N (frame == single-GPU frame), then bench.py at N.
# usage: bash tools/gpu_r2_multi.sh <tag> <N> [extra bench.py flags]
T=$1; N=$2; shift 2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_multi_check.py > gpurun_out/${T}_multi_n$N.log 2>&1; grep -E "MULTI_CHECK|False|Error" gpurun_out/${T}_multi_n$N.log | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err; tail -2 gpurun_out/${T}_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N", {k:round(d[k],2) for k in ("value","ms_per_step","fps","fps_at_2_in_flight","fps_latency")}, "e2e fps", round(d["e2e"]["fps"],1), "by rank", [round(x,1) for x in d["frame"]["timed_region_ms_by_rank"]], "cfg5 fps", d.get("config5",{}).get("fps"), d["frame"].get("partition"), d["frame"].get("bounds"), d["frame"].get("block_costs_by_round"))
PY
