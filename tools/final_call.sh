#!/bin/bash
# end-of-round measurement: bench line, ncu launch list of one generate(), ncu --set full of the decode kernels
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_generate_b128_maxlen4.csv python tools/profile_run.py --batch 128 --max-len 4 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'gemm_skinny_kernel|decode_attention_kernel|add_rmsnorm_row' -s 31710 -c 14 -f -o gpurun_out/decode_kernels_full \
    python tools/profile_run.py --batch 128 --max-len 160 --phase decode > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout 300 python tools/decode_timeline.py --out gpurun_out/decode_timeline_final.txt > /dev/null 2>> gpurun_out/tl.log
ls -la gpurun_out | tail -8
