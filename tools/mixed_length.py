"""SURVEY 8 row f3 on a mixed-length batch: B rows decode max_len steps, half of them are forced to emit the stop token
at step `--stop-at` (teacher forcing of the model's own ids otherwise, so the other rows are unchanged).  Finished rows
stop streaming their KV cache ("skip_finished") and, with "share_keys", their decode-attention CTAs take a share of the
live rows' keys; the same run with the skip off is the reference's behaviour (every row keeps decoding until all have
stopped).  `--stopped` = how many of the B rows stop.  Prints one JSON line with ms per decode step for the three modes."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--max-len", type=int, default=300)
ap.add_argument("--stop-at", type=int, default=8)
ap.add_argument("--stopped", type=int, default=-1, help="rows that stop at --stop-at (default: half)")
ap.add_argument("--policy", default="split24")
args = ap.parse_args()
B, L = args.batch, args.max_len
NS = B // 2 if args.stopped < 0 else args.stopped
eng = Engine(synth.synthetic_state_dict(), device=0, max_batch=B, max_new_tokens=L, policy=args.policy)
wave = synth.synthetic_waveforms(2 * B).cuda(); ids = synth.synthetic_prompt_ids(B).cuda()
s = torch.cuda.Stream()
res = {}
with torch.cuda.stream(s):
    def prefill():
        eng.encode(wave[:B], wave[B:]); eng.prefix(ids); eng.prefill(B, want_logits=False)
    prefill()
    own = eng.decode(B, L).clone()
    forced = own.clone()
    forced[:NS, args.stop_at] = 0                              # half of the rows emit the stop token at step stop_at
    # the forced matrix must be full-width for the kernel (row stride max_len)
    full = torch.zeros(B, L, dtype=torch.int32, device="cuda"); full[:, :forced.shape[1]] = forced
    for name, skip, share in (("share_keys", 1, 1), ("skip_finished", 1, 0), ("all_rows_keep_decoding", 0, 0)):
        eng.set_option("skip_finished", skip)
        eng.set_option("share_keys", share)
        prefill(); toks = eng.decode(B, L, forced_tokens=full)     # warm-up / graph capture
        prefill()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        toks = eng.decode(B, L, forced_tokens=full)
        e1.record(s)
        torch.cuda.synchronize()
        res[name] = {"decode_ms_per_step": e0.elapsed_time(e1) / toks.shape[1], "steps": int(toks.shape[1]),
                     "unfinished_rows_identical_to_unforced_run": bool(torch.equal(toks[NS:].cpu(), own[NS:, :toks.shape[1]].cpu())),
                     "unfinished_rows_ids_equal_frac": float((toks[NS:].cpu() == own[NS:, :toks.shape[1]].cpu()).float().mean())}
print(json.dumps({"batch": B, "max_len": L, "rows_stopped_at_step": {"rows": NS, "step": args.stop_at}, "policy": args.policy, **res}), flush=True)
eng.close()
