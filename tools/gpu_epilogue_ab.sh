#!/bin/bash
# GPU tests on the current tree, then prefill A/B of the plain GEMM epilogue (row-per-thread vs staged 128-byte rows).
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/gpu_epilogue_ab.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > gpurun_out/epi_tests.log; tail -5 gpurun_out/epi_tests.log
rm -f gpurun_out/prefill_ab_ref_*.pt
for v in 1 0 0; do
  timeout 200 python tools/prefill_ab.py --opt epilogue_rows=$v --tag epilogue_rows_$v 2>&1 | tail -1 | tee -a gpurun_out/prefill_ab_epilogue.jsonl
done
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/epi_bench.json 2> gpurun_out/epi_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/epi_bench.json")); p = d["phases"]
print("tokens/s", round(d["value"]), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]), "encoder ms", round(p["encoder_ms"], 1),
      "lm prefill ms", round(p["lm_prefill_ms"], 1), "decode ms/step", round(p["decode_ms_per_token_step"], 4))
PY
