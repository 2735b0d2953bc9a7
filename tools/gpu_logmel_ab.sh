#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > gpurun_out/logmel_tests.log; tail -8 gpurun_out/logmel_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -k regex:logmel -c 2 python tools/profile_run.py --batch 128 --max-len 4 --policy split24 --phase prefill 2>&1 | grep -E "gpu__time_duration|inst_executed|issue_active" | head -8
for i in 1; do timeout 200 python tools/prefill_ab.py --tag logmel_radix8 2>&1 | tail -1; done
