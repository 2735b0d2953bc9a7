#!/bin/bash
# usage: tools/gpu_multi.sh N "workload ..."   -- torchrun benches on N GPUs of one box (+ the 2-GPU wrapper test at N=2)
N=$1; shift
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 2>&1 | tail -5 > gpurun_out/r2_tests_2gpu.log; tail -3 gpurun_out/r2_tests_2gpu.log
fi
port=29520
for wl in "$@"; do
  port=$((port+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 2 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/r2_scale_${wl}_n${N}.json 2> gpurun_out/r2_scale_${wl}_n${N}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_scale_${wl}_n${N}.json").read().strip().splitlines()[-1]); p=d["phases"]
    print("${wl} N=${N}", "value",round(d["value"]),"ms",round(d["ms_per_step"],1),"e2e",round(d["e2e"]["value"]),"pairs/s",round(d["config"]["pairs_per_s"],1),"pairs_per_gpu",d["config"]["pairs_per_gpu"],"enc",round(p["encoder_ms"],1),"lm",round(p["lm_prefill_ms"],1),"dec",round(p["decode_ms_per_token_step"],4))
except Exception as e:
    print("${wl} N=${N} FAILED", e); print(open("gpurun_out/r2_scale_${wl}_n${N}.err").read()[-1500:])
PY
done
