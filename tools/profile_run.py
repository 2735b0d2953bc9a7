"""Workload for ncu: one warm generate(), then one profiled generate() between cudaProfilerStart/Stop.
   ncu --profile-from-start off ... python tools/profile_run.py --batch 128 --max-len 4"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--max-len", type=int, default=4)
ap.add_argument("--policy", default="split")
ap.add_argument("--phase", default="generate", choices=["generate", "decode", "prefill"])
args = ap.parse_args()
B = args.batch
eng = Engine(synth.synthetic_state_dict(), device=0, max_batch=B, max_new_tokens=max(args.max_len, 8), policy=args.policy)

wave = synth.synthetic_waveforms(2 * B).cuda(); ids = synth.synthetic_prompt_ids(B).cuda()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    eng.generate(wave[:B], wave[B:], ids, args.max_len)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    if args.phase == "generate":
        eng.generate(wave[:B], wave[B:], ids, args.max_len)
    elif args.phase == "prefill":
        eng.encode(wave[:B], wave[B:]); eng.prefix(ids); eng.prefill(B, want_logits=False)
    else:
        eng.decode(B, args.max_len)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("launches", eng.kernel_launches)
