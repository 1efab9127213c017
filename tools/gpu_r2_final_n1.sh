#!/bin/bash
# final single-GPU visit of the round: GPU tests, smoke, bench line (with the committed ncu facts of this kernel), reference arm, parity report
mkdir -p gpurun_out
T=r2_29
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo rc $? >> gpurun_out/${T}_pytest.log; grep -E "passed|failed|FAILED|^rc|^E  " gpurun_out/${T}_pytest.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -2 gpurun_out/${T}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 300 python tools/gpu_parity_report.py > gpurun_out/${T}_parity.log 2>&1; tail -25 gpurun_out/${T}_parity.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_29_bench_n1.json").read().strip().splitlines()[-1])
print({k:round(d[k],2) for k in ("value","ms_per_step","fps","fps_at_2_in_flight","fps_latency")}, "e2e", round(d["e2e"]["fps"],1), "cfg5", round(d["config5"]["fps"],1), "full", round(d["full_sampling"]["fps"],1), "hbm frac", d["roofline"]["frac"], "issue frac", d["roofline_fp32"]["frac"], d["cpu_baseline"]["value"], d["clocks"])
r=json.loads(open("gpurun_out/r2_29_bench_ref.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"]["cores"], r["ms_per_step"])
PY
