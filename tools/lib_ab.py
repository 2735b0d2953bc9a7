"""A/B of two BUILDS of the library inside one gpurun call (one process per build, alternate them in the shell):

    python tools/lib_ab.py --lib tools/_ab/libmellow_b200_<commit>.so --tag old
    python tools/lib_ab.py --tag new

Times the decode-attention kernel like bench.py's roofline leg (120 back-to-back launches, B = 128, ctx 539) and the
300-step decode loop; one JSON line.  Only options that exist in both builds may be used."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default="")
ap.add_argument("--tag", default="")
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--max-len", type=int, default=300)
ap.add_argument("--policy", default="split24")
args = ap.parse_args()
from mellow_b200 import native
if args.lib:
    native.LIB_PATH = os.path.abspath(args.lib)
    native.is_stale = lambda: False
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine

B, L = args.batch, args.max_len
eng = Engine(synth.synthetic_state_dict(), device=0, max_batch=B, max_new_tokens=L, policy=args.policy)
wave = synth.synthetic_waveforms(2 * B).cuda(); ids = synth.synthetic_prompt_ids(B).cuda()
s = torch.cuda.Stream()


def timed(fn, reps):
    best = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); fn(); e1.record(s)
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    return best


with torch.cuda.stream(s):
    eng.encode(wave[:B], wave[B:]); eng.prefix(ids); eng.prefill(B, want_logits=False)
    eng.decode(B, L)
    dec = [t / L for t in timed(lambda: eng.decode(B, L), 3)]
    eng.bench_decode_attention(B, 389 + L // 2, 120)
    att = [t / 120 * 1e3 for t in timed(lambda: eng.bench_decode_attention(B, 389 + L // 2, 120), 5)]
print(json.dumps({"tag": args.tag, "lib": os.path.basename(native.LIB_PATH), "decode_ms_per_step": dec, "attention_us_per_launch": att}), flush=True)
eng.close()
