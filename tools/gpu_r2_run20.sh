#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_parity_report.py 256 144 2>&1 | cut -c1-150 > gpurun_out/r2_20_parity.log; cat gpurun_out/r2_20_parity.log
timeout 300 python tools/gpu_frame_probe.py --tag step --slots 1 2 4 --blocks 3 4 5 2>&1 | tee gpurun_out/r2_20_probe.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_20_pytest.log 2>&1; echo rc $? >> gpurun_out/r2_20_pytest.log; grep -E "passed|failed|FAILED|^rc|^E  " gpurun_out/r2_20_pytest.log | head -20
