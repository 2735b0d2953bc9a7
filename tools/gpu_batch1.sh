#!/bin/bash
# round 2, GPU batch 1: correctness after the clean-up + first A/B numbers (L2 prefetch, 24-bit KV)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -60 > gpurun_out/r2_tests1.log
tail -5 gpurun_out/r2_tests1.log
rm -f gpurun_out/decode_ab_ref_*.pt gpurun_out/r2_ab1.jsonl
for cfg in "split kv_prefetch=0" "split kv_prefetch=-1" "split24 kv_prefetch=0" "split24 kv_prefetch=-1" "split24 kv_prefetch=389"; do
  set -- $cfg
  python tools/decode_ab.py --policy $1 --opt $2 --tag "$1_$2" >> gpurun_out/r2_ab1.jsonl 2>> gpurun_out/r2_ab1.err
done
cat gpurun_out/r2_ab1.jsonl
python tools/decode_timeline.py --policy split24 --out gpurun_out/r2_timeline_split24_pf.txt > /dev/null 2>&1
python tools/decode_timeline.py --policy split24 --opt kv_prefetch=0 --out gpurun_out/r2_timeline_split24_nopf.txt > /dev/null 2>&1
tail -12 gpurun_out/r2_timeline_split24_pf.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_attention -c 3 -o gpurun_out/r2_attn24 python tools/profile_attn.py split24 > gpurun_out/r2_ncu_attn24.log 2>&1
python bench.py --steps 3 --warmup 3 --policy split24 > gpurun_out/r2_bench1_split24.json 2> gpurun_out/r2_bench1.err
cat gpurun_out/r2_bench1_split24.json
