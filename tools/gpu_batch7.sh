#!/bin/bash
# round 2, GPU batch 7: attention softmax micro-optimisations + cheaper GELU epilogue; bench with robust phases
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > gpurun_out/r2_tests7.log
tail -6 gpurun_out/r2_tests7.log | cut -c1-300
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench7.json")); p=d["phases"]
print("value",round(d["value"]),"ms",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"]),"enc",round(p["encoder_ms"],1),"lm",round(p["lm_prefill_ms"],1),"dec",round(p["decode_ms_per_token_step"],4),"roof",round(d["roofline"]["frac"],3),d["roofline"]["us_per_launch"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_prefill_b128_v7.csv python tools/profile_run.py --batch 128 --phase prefill --policy split24 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r2_launches_prefill_b128_v7.csv 16
