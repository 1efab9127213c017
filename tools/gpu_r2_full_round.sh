#!/bin/bash
# final-candidate kernel: probe, tests, bench, ncu launch list + full capture
mkdir -p gpurun_out
timeout 300 python tools/gpu_frame_probe.py --tag final --slots 1 2 4 --blocks 3 2>&1 | tee gpurun_out/r2_25_probe.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_25_pytest.log 2>&1; echo rc $? >> gpurun_out/r2_25_pytest.log; grep -E "passed|failed|FAILED|^rc|^E  " gpurun_out/r2_25_pytest.log | head
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_25_bench_n1.json 2> gpurun_out/r2_25_bench_n1.err; tail -2 gpurun_out/r2_25_bench_n1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2_25_bench_ref.json 2> gpurun_out/r2_25_bench_ref.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_25_bench_n1.json").read().strip().splitlines()[-1])
print({k:round(d[k],2) for k in ("value","ms_per_step","fps","fps_at_2_in_flight","fps_latency")}, d["frame"]["rank0_stage_ms_one_frame_alone"], "e2e", round(d["e2e"]["fps"],1), "cfg5", round(d["config5"]["fps"],1), "full", round(d["full_sampling"]["fps"],1), d["roofline"]["frac"], d["roofline_fp32"]["frac"], d["cpu_baseline"])
r=json.loads(open("gpurun_out/r2_25_bench_ref.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r["cpu_baseline"]["cores"], r["ms_per_step"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_25_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-measurements > gpurun_out/r2_25_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 3 -c 3 -o gpurun_out/r2_25_prof_integrate python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-side-measurements > gpurun_out/r2_25_ncu_full.log 2>&1
python tools/ncu_step_mix.py gpurun_out/r2_25_prof_integrate.ncu-rep gpurun_out/r2_25_ncu_traffic.json > gpurun_out/r2_25_step_mix.log 2>&1; grep -E "sass_instr|sass_fp|issue_active|pipe_fma|warp_instr|lanes" gpurun_out/r2_25_step_mix.log
python tools/summarise_ncu.py gpurun_out/r2_25_prof_integrate.ncu-rep > gpurun_out/r2_25_ncu_k_integrate.txt 2>&1
python tools/summarise_launches.py gpurun_out/r2_25_launches.csv 2>&1 | tail -15
