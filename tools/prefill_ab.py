"""A/B timing of the prefill (front end + encoder, prefix, LM prefill) under the library's options (mb_set_option):

    python tools/prefill_ab.py --batch 128 --opt epilogue_rows=1 --tag row_epilogue

Prints one JSON line: encoder ms, LM prefill ms (CUDA events on the launch stream, 2 warm-up passes, mean of --iters),
the max |difference| of the prefill logits against the first configuration run in this gpurun call (kept in
gpurun_out/) and whether their argmax agrees."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--policy", default="split24")
ap.add_argument("--tag", default="")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--opt", action="append", default=[], help="name=value for mb_set_option")
args = ap.parse_args()
B = args.batch
eng = Engine(synth.synthetic_state_dict(), device=0, max_batch=B, max_new_tokens=8, policy=args.policy)
for kv in args.opt:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
wave = synth.synthetic_waveforms(2 * B).cuda(); ids = synth.synthetic_prompt_ids(B).cuda()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        eng.encode(wave[:B], wave[B:]); eng.prefix(ids); logits = eng.prefill(B)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.iters + 1)]
    e[0].record(s)
    for i in range(args.iters):
        eng.encode(wave[:B], wave[B:]); eng.prefix(ids)
        e[2 * i + 1].record(s)
        eng.prefill(B, want_logits=False)
        e[2 * i + 2].record(s)
    torch.cuda.synchronize()
enc = sum(e[2 * i].elapsed_time(e[2 * i + 1]) for i in range(args.iters)) / args.iters
lm = sum(e[2 * i + 1].elapsed_time(e[2 * i + 2]) for i in range(args.iters)) / args.iters
os.makedirs("gpurun_out", exist_ok=True)
ref_path = os.path.join("gpurun_out", f"prefill_ab_ref_b{B}_{args.policy}.pt")
diff, same = None, None
if os.path.isfile(ref_path):
    ref = torch.load(ref_path)
    diff = float((ref - logits.cpu()).abs().max())
    same = bool(torch.equal(ref.argmax(-1), logits.cpu().argmax(-1)))
else:
    torch.save(logits.cpu(), ref_path)
print(json.dumps({"tag": args.tag, "policy": args.policy, "opts": args.opt, "batch": B, "encoder_ms": enc, "lm_prefill_ms": lm,
                  "prefill_ms": enc + lm, "finite": bool(torch.isfinite(logits).all()),
                  "max_abs_logit_diff_vs_first_config": diff, "argmax_equal_first_config": same}), flush=True)
eng.close()
