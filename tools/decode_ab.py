"""A/B timing of the decode loop under the library's options (mb_set_option) / environment switches:

    python tools/decode_ab.py --batch 128 --max-len 300 --policy split24 --opt kv_prefetch=0 --tag no_prefetch

Prints one JSON line: decode ms per token step (CUDA events on the launch stream, 1 warm-up decode), prefill ms, and
whether the greedy ids equal those of the first configuration run in this gpurun call (kept in gpurun_out/)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--max-len", type=int, default=300)
ap.add_argument("--policy", default="split")
ap.add_argument("--tag", default="")
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--opt", action="append", default=[], help="name=value for mb_set_option")
args = ap.parse_args()
B = args.batch
eng = Engine(synth.synthetic_state_dict(), device=0, max_batch=B, max_new_tokens=args.max_len, policy=args.policy)
for kv in args.opt:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
wave = synth.synthetic_waveforms(2 * B).cuda(); ids = synth.synthetic_prompt_ids(B).cuda()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    def prefill():
        eng.encode(wave[:B], wave[B:]); eng.prefix(ids); eng.prefill(B, want_logits=False)
    prefill()
    toks = eng.decode(B, args.max_len)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(s)
    prefill()
    e[1].record(s)
    for _ in range(args.iters):
        toks = eng.decode(B, args.max_len)
    e[2].record(s)
    torch.cuda.synchronize()
ref_path = os.path.join("gpurun_out", f"decode_ab_ref_b{B}_l{args.max_len}_{args.policy}.pt")
same, diff_rows, first_diff = None, None, None
os.makedirs("gpurun_out", exist_ok=True)
if os.path.isfile(ref_path):
    ref = torch.load(ref_path)
    same = bool(torch.equal(ref, toks.cpu()))
    ne = (ref != toks.cpu())
    diff_rows = int(ne.any(dim=1).sum())
    first_diff = int(ne.any(dim=0).nonzero()[0]) if diff_rows else None
else:
    torch.save(toks.cpu(), ref_path)
print(json.dumps({"tag": args.tag, "policy": args.policy, "opts": args.opt, "env": {k: v for k, v in os.environ.items() if k.startswith("MB_")}, "batch": B,
                  "steps": int(toks.shape[1]), "prefill_ms": e[0].elapsed_time(e[1]),
                  "decode_ms_per_step": e[1].elapsed_time(e[2]) / args.iters / toks.shape[1],
                  "tokens_equal_first_config": same, "rows_differing": diff_rows, "first_differing_step": first_diff}), flush=True)
eng.close()
