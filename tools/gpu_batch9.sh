#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/decode_ab_ref_*.pt gpurun_out/r2_ab9.jsonl
for w in 0 1 2 3; do
  timeout 200 python tools/decode_ab.py --policy split24 --opt decode_wide=$w --tag "decode_wide=$w" >> gpurun_out/r2_ab9.jsonl 2>> gpurun_out/r2_ab9.err
done
cut -c1-330 gpurun_out/r2_ab9.jsonl
timeout 200 python tools/decode_timeline.py --policy split24 --opt decode_wide=3 --out gpurun_out/r2_timeline_wide3.txt > /dev/null 2>&1
tail -12 gpurun_out/r2_timeline_wide3.txt
timeout 200 python bench.py --steps 2 --warmup 3 --workload configs4 --no-cpu-baseline > gpurun_out/r2_scale_configs4_n1.json 2> gpurun_out/r2_bench9.err; cut -c1-400 gpurun_out/r2_scale_configs4_n1.json
