#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "tails or variants" 2>&1 | tail -4
rm -f gpurun_out/decode_ab_ref_*.pt gpurun_out/r2_ab11.jsonl
for cfg in "tail_sync=0" "tail_sync=1" "tail_sync=0" "tail_sync=1"; do
  timeout 200 python tools/decode_ab.py --policy split24 --opt $cfg --tag "$cfg" >> gpurun_out/r2_ab11.jsonl 2>> gpurun_out/r2_ab11.err
done
cut -c1-300 gpurun_out/r2_ab11.jsonl; tail -3 gpurun_out/r2_ab11.err
timeout 200 python tools/decode_timeline.py --policy split24 --opt tail_sync=0 --out gpurun_out/r2_timeline_sync0.txt > /dev/null 2>&1
tail -13 gpurun_out/r2_timeline_sync0.txt | head -6
timeout 200 python tools/decode_timeline.py --policy split24 --opt tail_sync=1 --out gpurun_out/r2_timeline_sync1.txt > /dev/null 2>&1
tail -13 gpurun_out/r2_timeline_sync1.txt
