#!/bin/bash
# leaner step (292-299 SASS instructions): probe, parity tests, bench
mkdir -p gpurun_out
T=r2_24
timeout 300 python tools/gpu_frame_probe.py --tag lean --slots 1 2 4 --blocks 3 2>&1 | tee gpurun_out/${T}_probe.log
timeout 200 python tools/gpu_frame_probe.py --tag cfg5 --metric double_unequal_kerr --slots 1 4 --blocks 3 2>&1 | tee -a gpurun_out/${T}_probe.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo rc $? >> gpurun_out/${T}_pytest.log; grep -E "passed|failed|FAILED|^rc|^E  " gpurun_out/${T}_pytest.log | head -20
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -2 gpurun_out/${T}_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_24_bench_n1.json").read().strip().splitlines()[-1])
print({k:round(d[k],2) for k in ("value","ms_per_step","fps","fps_at_2_in_flight","fps_latency")}, d["frame"]["rank0_stage_ms_one_frame_alone"], "e2e", round(d["e2e"]["fps"],1), "cfg5", round(d["config5"]["fps"],1), "full", round(d["full_sampling"]["fps"],1))
PY
