#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/prefill_ab_ref_*.pt
for v in 1 2 1 2 0; do
  timeout 120 python tools/prefill_ab.py --opt cta_pairs=$v --tag cta_pairs_$v 2>&1 | tail -1 | cut -c1-330 | tee -a gpurun_out/prefill_ab_pairs.jsonl
done
