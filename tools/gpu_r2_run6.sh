#!/bin/bash
# round 2, GPU call 6: tests + bench after the value-based merge and the loop changes; ncu capture of k_integrate; FP32 pipe microbench under ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_6_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_6_pytest.log
grep -E "passed|failed|FAILED|rc" gpurun_out/r2_6_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_6_bench_n1.json 2> gpurun_out/r2_6_bench_n1.err; tail -3 gpurun_out/r2_6_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_6_bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","fps","fps_at_2_in_flight","fps_latency")}, d["frame"]["rank0_stage_ms_one_frame_alone"], d["e2e"]["fps"], d.get("config5",{}).get("fps"), d.get("full_sampling",{}).get("fps"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 3 -c 3 -o gpurun_out/r2_6_prof_integrate python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-side-measurements > gpurun_out/r2_6_ncu_full.log 2>&1
python tools/ncu_step_mix.py gpurun_out/r2_6_prof_integrate.ncu-rep gpurun_out/r2_6_ncu_traffic.json > gpurun_out/r2_6_step_mix.log 2>&1; tail -45 gpurun_out/r2_6_step_mix.log
python tools/summarise_ncu.py gpurun_out/r2_6_prof_integrate.ncu-rep > gpurun_out/r2_6_ncu_k_integrate.txt 2>&1
nvcc -O3 -DITERS=262144 -gencode arch=compute_100a,code=sm_100a -lineinfo -o /tmp/fp32_long tools/microbench/fp32_pipes.cu && /tmp/fp32_long > gpurun_out/r2_6_fp32_long.log 2>&1; tail -8 gpurun_out/r2_6_fp32_long.log
timeout 600 ncu --metrics sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -s 30 -c 10 --csv --log-file gpurun_out/r2_6_fp32_ncu.csv /tmp/fp32_long > /dev/null 2>&1
cut -d, -f5,13- gpurun_out/r2_6_fp32_ncu.csv | head -60
