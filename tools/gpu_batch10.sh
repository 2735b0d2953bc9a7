#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/decode_ab_ref_*.pt gpurun_out/r2_ab10.jsonl
for cfg in "pf_partner=0 tails_mode=1" "pf_partner=1 tails_mode=1" "pf_partner=0 tails_mode=2" "pf_partner=1 tails_mode=2"; do
  set -- $cfg
  timeout 200 python tools/decode_ab.py --policy split24 --opt $1 --opt $2 --tag "$1_$2" >> gpurun_out/r2_ab10.jsonl 2>> gpurun_out/r2_ab10.err
done
cut -c1-330 gpurun_out/r2_ab10.jsonl; tail -3 gpurun_out/r2_ab10.err
timeout 200 python tools/decode_timeline.py --policy split24 --opt pf_partner=1 --opt tails_mode=2 --out gpurun_out/r2_timeline_pf_tm2.txt > /dev/null 2>&1
tail -14 gpurun_out/r2_timeline_pf_tm2.txt
timeout 300 python -m pytest tests/test_gpu_long_parity.py -m gpu -q -s --timeout 180 2>&1 | grep -E "teacher-forced|passed|failed" > gpurun_out/r2_long_parity_report.txt; cat gpurun_out/r2_long_parity_report.txt | cut -c1-400
