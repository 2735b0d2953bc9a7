#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > gpurun_out/r2_tests8.log
tail -6 gpurun_out/r2_tests8.log | cut -c1-300
timeout 200 python tools/decode_timeline.py --policy split24 --out gpurun_out/r2_timeline_tails64.txt > /dev/null 2>&1
tail -16 gpurun_out/r2_timeline_tails64.txt
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench8.json")); p=d["phases"]
print("value",round(d["value"]),"ms",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"]),"enc",round(p["encoder_ms"],1),"lm",round(p["lm_prefill_ms"],1),"dec",round(p["decode_ms_per_token_step"],4),"roof",round(d["roofline"]["frac"],3),d["roofline"]["us_per_launch"])
PY
