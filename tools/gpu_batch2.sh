#!/bin/bash
# round 2, GPU batch 2: warp-autonomous decode attention, cluster tails, encoder captures, fast-policy root cause
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests2.log
tail -6 gpurun_out/r2_tests2.log
rm -f gpurun_out/decode_ab_ref_*.pt gpurun_out/r2_ab2.jsonl
for cfg in "attn_variant=0 decode_tails=0" "attn_variant=1 decode_tails=0" "attn_variant=1 decode_tails=1" "attn_variant=0 decode_tails=1"; do
  set -- $cfg
  python tools/decode_ab.py --policy split24 --opt $1 --opt $2 --tag "$1_$2" >> gpurun_out/r2_ab2.jsonl 2>> gpurun_out/r2_ab2.err
done
python tools/decode_ab.py --policy split --opt attn_variant=1 --tag "split_v1" >> gpurun_out/r2_ab2.jsonl 2>> gpurun_out/r2_ab2.err
cat gpurun_out/r2_ab2.jsonl
python tools/decode_timeline.py --policy split24 --out gpurun_out/r2_timeline_v1.txt > /dev/null 2>&1
python tools/decode_timeline.py --policy split24 --opt decode_tails=1 --out gpurun_out/r2_timeline_v1_tails.txt > /dev/null 2>&1
tail -12 gpurun_out/r2_timeline_v1.txt; tail -12 gpurun_out/r2_timeline_v1_tails.txt
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_attention -c 3 -o gpurun_out/r2_attn_warp24 python tools/profile_attn.py split24 > gpurun_out/r2_ncu_attn_warp24.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:logmel|window_attention|prefill_attention|patch_embed" -c 8 -o gpurun_out/r2_encoder_kernels python tools/profile_run.py --batch 128 --phase prefill --policy split24 > gpurun_out/r2_ncu_encoder.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_umma -s 4 -c 8 -o gpurun_out/r2_encoder_gemms python tools/profile_run.py --batch 128 --phase prefill --policy split24 > gpurun_out/r2_ncu_encoder_gemms.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_generate_b128_maxlen4.csv python tools/profile_run.py --batch 128 --max-len 4 --policy split24 > /dev/null 2>&1
for i in 1 2 3 4 5 6; do timeout 120 python tools/fast_repro.py 4 >> gpurun_out/r2_fast_repro.jsonl 2>> gpurun_out/r2_fast_repro.err; done
cat gpurun_out/r2_fast_repro.jsonl
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python tools/fast_repro.py 2 > gpurun_out/r2_sanitizer_racecheck_fast.log 2>&1
tail -5 gpurun_out/r2_sanitizer_racecheck_fast.log
timeout 500 compute-sanitizer --tool initcheck --print-limit 20 python tools/fast_repro.py 2 > gpurun_out/r2_sanitizer_initcheck_fast.log 2>&1
tail -5 gpurun_out/r2_sanitizer_initcheck_fast.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
cat gpurun_out/r2_bench2.json
