#!/bin/bash
# one gpurun call: decode A/B over row-group configurations, then the GPU test-suite
mkdir -p gpurun_out
rm -f gpurun_out/decode_ab_ref_*.pt gpurun_out/ab1.jsonl
run() { echo "== $*" >> gpurun_out/ab1.log; env "$@" timeout 300 python tools/decode_ab.py --tag "$*" >> gpurun_out/ab1.jsonl 2>> gpurun_out/ab1.log; }
run MB_DECODE_GROUPS=1 MB_DECODE_COMPACT=0
run MB_DECODE_GROUPS=2 MB_DECODE_COMPACT=1
run MB_DECODE_GROUPS=1 MB_DECODE_COMPACT=1
run MB_DECODE_GROUPS=2 MB_DECODE_COMPACT=0
run MB_DECODE_GROUPS=3 MB_DECODE_COMPACT=1
run MB_DECODE_GROUPS=4 MB_DECODE_COMPACT=1
cat gpurun_out/ab1.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t1.log 2>&1; tail -5 gpurun_out/t1.log
