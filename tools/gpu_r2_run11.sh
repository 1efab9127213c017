#!/bin/bash
mkdir -p gpurun_out
nproc
for V in 0 1; do
GRT_COMM_DIRECT_WRITE=$V timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$V bench.py --gpus 2 --steps 20 --warmup 3 --no-side-measurements > gpurun_out/r2_11_bench_v$V.json 2> gpurun_out/r2_11_bench_v$V.err; tail -2 gpurun_out/r2_11_bench_v$V.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_11_bench_v$V.json").read().strip().splitlines()[-1])
print("direct=$V N=2", {k:round(d[k],2) for k in ("ms_per_step","fps","fps_at_2_in_flight","fps_latency")}, "by rank", d["frame"]["timed_region_ms_by_rank"], "host ms/submit", d["frame"]["host_ms_per_submit_rank0"], "e2e fps", round(d["e2e"]["fps"],1))
PY
done
for R in 0 1; do
  GRT_BENCH_EMULATE_WORLD=2 GRT_BENCH_EMULATE_RANK=$R GRT_BENCH_FRAMES_IN_FLIGHT=8 GRT_BENCH_BLOCKS_PER_SM=2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-side-measurements 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('emulated rank $R of 2: ms_per_step %.2f host ms/submit %.3f' % (d['ms_per_step'], d['frame']['host_ms_per_submit_rank0']))
"
done
