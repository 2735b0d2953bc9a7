#!/bin/bash
# GPU tests, prefill timing and the per-launch time of the tcgen05 prefill attention kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -30 > gpurun_out/attn_tests.log; tail -5 gpurun_out/attn_tests.log
rm -f gpurun_out/prefill_ab_ref_*.pt
for i in 1 2; do timeout 200 python tools/prefill_ab.py --tag attention_v2 2>&1 | tail -1 | tee -a gpurun_out/prefill_ab_attn.jsonl; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:prefill_attention -c 6 python tools/profile_run.py --batch 128 --max-len 4 --policy split24 --phase prefill 2>&1 | grep -E "gpu__time_duration|prefill_attention" | head -14
