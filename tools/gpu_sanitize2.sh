#!/bin/bash
# compute-sanitizer on the kernels rewritten late in round 2: the encoder pass (radix-8 log-mel, staged GEMM epilogues) under
# memcheck, and racecheck restricted to the kernels that stage data through shared memory with warp-level synchronisation only
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python tools/check_frontend.py > gpurun_out/r2_sanitizer_encoder_memcheck.log 2>&1
echo "== encoder memcheck rc=$?"; grep -E "ERROR SUMMARY|rows_finite|Invalid" gpurun_out/r2_sanitizer_encoder_memcheck.log | head -5 | cut -c1-300
timeout 240 compute-sanitizer --tool racecheck --kernel-regex kns=logmel --print-limit 10 python tools/check_frontend.py > gpurun_out/r2_sanitizer_logmel_racecheck.log 2>&1
echo "== logmel racecheck rc=$?"; grep -E "RACECHECK SUMMARY|rows_finite|hazard" gpurun_out/r2_sanitizer_logmel_racecheck.log | head -5 | cut -c1-300
timeout 300 compute-sanitizer --tool racecheck --kernel-regex kns=prefill_attention --print-limit 10 python tools/check_options.py > gpurun_out/r2_sanitizer_prefill_attention_racecheck.log 2>&1
echo "== prefill attention racecheck rc=$?"; grep -E "RACECHECK SUMMARY|prefill_max_err|hazard" gpurun_out/r2_sanitizer_prefill_attention_racecheck.log | head -5 | cut -c1-300
timeout 300 compute-sanitizer --tool racecheck --kernel-regex kns=gemm_umma --print-limit 10 python tools/check_frontend.py > gpurun_out/r2_sanitizer_gemm_umma_encoder_racecheck.log 2>&1
echo "== persistent GEMM (encoder pass) racecheck rc=$?"; grep -E "RACECHECK SUMMARY|rows_finite|hazard" gpurun_out/r2_sanitizer_gemm_umma_encoder_racecheck.log | head -5 | cut -c1-300
timeout 300 compute-sanitizer --tool racecheck --kernel-regex kns=gemm_umma --print-limit 10 python tools/check_options.py > gpurun_out/r2_sanitizer_gemm_umma_lm_racecheck.log 2>&1
echo "== persistent GEMM (LM prefill) racecheck rc=$?"; grep -E "RACECHECK SUMMARY|prefill_max_err|hazard" gpurun_out/r2_sanitizer_gemm_umma_lm_racecheck.log | head -5 | cut -c1-300
