#!/bin/bash
# compute-sanitizer over the default path at B = 2 (prefill with the tcgen05 attention, 8 decode steps with the cluster tails
# and the bulk-copy attention): memcheck, racecheck, synccheck, initcheck.  Every run bounded.
mkdir -p gpurun_out
for tool in memcheck synccheck initcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 10 python tools/check_options.py > gpurun_out/r2_sanitizer_default_path_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|prefill_max_err|Invalid|hazard|Uninitialized|Barrier error" gpurun_out/r2_sanitizer_default_path_$tool.log | head -8 | cut -c1-300
done
