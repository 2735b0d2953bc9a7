#!/bin/bash
# end-of-round measurement: GPU tests, bench line, ncu launch list of one generate(), decode timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests_final.log 2>&1; tail -2 gpurun_out/tests_final.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; wc -l gpurun_out/bench_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_generate_b128_maxlen4.csv python tools/profile_run.py --batch 128 --max-len 4 > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log
timeout 300 python tools/decode_timeline.py --out gpurun_out/decode_timeline_final.txt > /dev/null 2>> gpurun_out/tl.log
