#!/bin/bash
# round 2, GPU batch 5: tcgen05 attention after the divergence fix + option matrix; every step bounded
mkdir -p gpurun_out
timeout 120 python tools/debug_attn.py prefill_attn=1 > gpurun_out/dbg_attn1.log 2>&1; echo "attn1 rc=$?"; tail -2 gpurun_out/dbg_attn1.log | cut -c1-500
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -40 > gpurun_out/r2_tests5.log
tail -15 gpurun_out/r2_tests5.log | cut -c1-300
rm -f gpurun_out/decode_ab_ref_*.pt gpurun_out/r2_ab5.jsonl
for cfg in "attn_variant=1 decode_tails=0 decode_cluster=0" "attn_variant=2 decode_tails=0 decode_cluster=0" "attn_variant=1 decode_tails=1 decode_cluster=0" "attn_variant=1 decode_tails=1 decode_cluster=1" "attn_variant=2 decode_tails=1 decode_cluster=1" "attn_variant=1 decode_tails=0 decode_cluster=1"; do
  set -- $cfg
  timeout 200 python tools/decode_ab.py --policy split24 --opt $1 --opt $2 --opt $3 --tag "$1_$2_$3" >> gpurun_out/r2_ab5.jsonl 2>> gpurun_out/r2_ab5.err
done
cat gpurun_out/r2_ab5.jsonl | cut -c1-420
timeout 200 python tools/decode_timeline.py --policy split24 --opt decode_tails=1 --opt decode_cluster=1 --out gpurun_out/r2_timeline_tails_cluster.txt > /dev/null 2>&1
tail -11 gpurun_out/r2_timeline_tails_cluster.txt
timeout 200 python tools/mixed_length.py > gpurun_out/r2_mixed_length.json 2> gpurun_out/r2_mixed_length.err; cat gpurun_out/r2_mixed_length.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_prefill_b128_umma_attn.csv python tools/profile_run.py --batch 128 --phase prefill --policy split24 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_prefill_b128_umma_attn.csv | head -12
