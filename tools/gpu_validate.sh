#!/bin/bash
# One bounded validation pass on a 1-GPU box (every step under its own timeout; results in gpurun_out/):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh'
# GPU parity tests, the smoke test, the bench line, the decode timeline, the ncu captures the bench / DESIGN cite
# (decode attention; one LM prefill layer: QKV, attention, o_proj, gate/up, down) and the launch list of a generate().
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > gpurun_out/validate_tests.log; tail -5 gpurun_out/validate_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/validate_bench.json")); p = d["phases"]
print("tokens/s", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), "encoder ms", round(p["encoder_ms"], 1),
      "lm prefill ms", round(p["lm_prefill_ms"], 1), "decode ms/step", round(p["decode_ms_per_token_step"], 4), "attention frac", round(d["roofline"]["frac"], 3))
PY
timeout 200 python tools/decode_timeline.py --policy split24 --out gpurun_out/validate_timeline.txt > /dev/null 2>&1; tail -9 gpurun_out/validate_timeline.txt
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_attention -c 3 -o gpurun_out/validate_decode_attention python tools/profile_attn.py split24 > /dev/null 2>&1
# layer 1 of the LM prefill: 55 encoder GEMMs + the 5 kernels of layer 0 are skipped
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:gemm_umma_kernel|prefill_attention_umma" --launch-skip 60 --launch-count 5 -o gpurun_out/validate_lm_prefill_layer python tools/profile_run.py --batch 128 --policy split24 --phase prefill > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/validate_launches_generate_b128_maxlen4.csv python tools/profile_run.py --batch 128 --max-len 4 --policy split24 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/validate_launches_generate_b128_maxlen4.csv 12
ls -la gpurun_out/*.ncu-rep
