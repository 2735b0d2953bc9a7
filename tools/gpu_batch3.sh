#!/bin/bash
# round 2, GPU batch 3: tcgen05 prefill attention, deferred-rstd fix for the cluster tails, mixed-length batch
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2_tests3.log
tail -25 gpurun_out/r2_tests3.log
rm -f gpurun_out/decode_ab_ref_*.pt gpurun_out/r2_ab3.jsonl
for cfg in "prefill_attn=0 decode_tails=0" "prefill_attn=1 decode_tails=1"; do
  set -- $cfg
  python tools/decode_ab.py --policy split24 --opt $1 --opt $2 --tag "$1_$2" >> gpurun_out/r2_ab3.jsonl 2>> gpurun_out/r2_ab3.err
done
cat gpurun_out/r2_ab3.jsonl; tail -3 gpurun_out/r2_ab3.err
python tools/decode_timeline.py --policy split24 --opt decode_tails=1 --out gpurun_out/r2_timeline_v1_tails_fix.txt > /dev/null 2>&1
tail -11 gpurun_out/r2_timeline_v1_tails_fix.txt
python tools/mixed_length.py > gpurun_out/r2_mixed_length.json 2> gpurun_out/r2_mixed_length.err; cat gpurun_out/r2_mixed_length.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_prefill_b128_umma_attn.csv python tools/profile_run.py --batch 128 --phase prefill --policy split24 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_prefill_b128_umma_attn.csv | head -12
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:prefill_attention_umma -c 2 -o gpurun_out/r2_prefill_attn_umma python tools/profile_run.py --batch 128 --phase prefill --policy split24 > gpurun_out/r2_ncu_prefill_attn.log 2>&1
