#!/bin/bash
# one gpurun call: decode A/B over the configurations listed in $1 (one "VAR=val VAR=val" set per line), then the GPU tests
mkdir -p gpurun_out
OUT=gpurun_out/${2:-ab}.jsonl
rm -f gpurun_out/decode_ab_ref_*.pt $OUT
while read -r line; do
  [ -z "$line" ] && continue
  echo "== $line" >> gpurun_out/ab.log
  env $line timeout 300 python tools/decode_ab.py --tag "$line" >> $OUT 2>> gpurun_out/ab.log
done < "$1"
cat $OUT
if [ -z "$NO_TESTS" ]; then timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; tail -5 gpurun_out/tests.log; fi
