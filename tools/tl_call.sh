#!/bin/bash
mkdir -p gpurun_out
bash tools/ab_call.sh tools/ab_cfg.txt ab7
timeout 300 python tools/decode_timeline.py --out gpurun_out/timeline_v3.txt > /dev/null 2>> gpurun_out/tl.log
head -1 gpurun_out/timeline_v3.txt; tail -16 gpurun_out/timeline_v3.txt
tail -3 gpurun_out/tl.log
