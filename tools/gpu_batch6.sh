#!/bin/bash
# round 2, GPU batch 6: defaults (bulk warp attention, cluster tails, tcgen05 prefill attention with exp2), bench + captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > gpurun_out/r2_tests6.log
tail -8 gpurun_out/r2_tests6.log | cut -c1-300
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; cat gpurun_out/r2_bench6.json | cut -c1-3000
timeout 200 python tools/decode_timeline.py --policy split24 --out gpurun_out/r2_timeline_defaults.txt > /dev/null 2>&1
tail -11 gpurun_out/r2_timeline_defaults.txt
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_attention -c 3 -o gpurun_out/r2_attn_bulk24 python tools/profile_attn.py split24 > gpurun_out/r2_ncu_attn_bulk24.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:prefill_attention_umma -c 2 -o gpurun_out/r2_prefill_attn_umma python tools/profile_run.py --batch 128 --phase prefill --policy split24 > gpurun_out/r2_ncu_prefill_attn.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_generate_b128_maxlen4_final.csv python tools/profile_run.py --batch 128 --max-len 4 --policy split24 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_generate_b128_maxlen4_final.csv 14
timeout 200 python bench.py --steps 2 --warmup 3 --workload configs1 --no-cpu-baseline > gpurun_out/r2_bench6_configs1.json 2>> gpurun_out/r2_bench6.err; cut -c1-600 gpurun_out/r2_bench6_configs1.json
timeout 300 python bench.py --steps 2 --warmup 3 --workload configs3 --no-cpu-baseline > gpurun_out/r2_bench6_configs3_n1.json 2>> gpurun_out/r2_bench6.err; cut -c1-600 gpurun_out/r2_bench6_configs3_n1.json
