#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests_final.log 2>&1; tail -3 gpurun_out/tests_final.log; grep -n "Error\|assert " gpurun_out/tests_final.log | head -12
timeout 300 python tools/decode_timeline.py --out gpurun_out/decode_timeline_final.txt > /dev/null 2>> gpurun_out/tl.log; tail -10 gpurun_out/decode_timeline_final.txt
