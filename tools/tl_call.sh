#!/bin/bash
mkdir -p gpurun_out
for pol in split split24; do
timeout 600 python bench.py --steps 2 --warmup 3 --policy $pol --no-cpu-baseline > gpurun_out/bench_$pol.json 2>/dev/null; python - <<PY
import json
d=json.load(open('gpurun_out/bench_$pol.json'))
print('$pol', round(d['value']), round(d['ms_per_step'],1), 'decode ms/step', round(d['phases']['decode_ms_per_token_step'],4), 'attn us', round(d['roofline']['us_per_launch'],2), 'frac', round(d['roofline']['frac'],3))
PY
done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
