#!/bin/bash
mkdir -p gpurun_out
bash tools/ab_call.sh tools/ab_cfg.txt ab13
timeout 300 python tools/decode_timeline.py --out gpurun_out/timeline_v11.txt > /dev/null 2>> gpurun_out/tl.log
head -1 gpurun_out/timeline_v11.txt; tail -18 gpurun_out/timeline_v11.txt | cut -c1-150
tail -3 gpurun_out/tl.log; grep -n "FAIL\|Error\|error" gpurun_out/tests.log | head
