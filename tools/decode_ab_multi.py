"""A/B timing of the decode loop over SEVERAL option sets inside one process (one engine, one prefill):

    python tools/decode_ab_multi.py --batch 128 --max-len 300 --policy split24 \
        --cfg base: --cfg d648:down_tail=648 --cfg o448:o_tail=448 --rounds 3

Every configuration is `tag:name=value,name=value` (mb_set_option; options a configuration does not name are reset to
the values given by --defaults).  The configurations are timed round-robin `--rounds` times so that clock / thermal
drift hits all of them alike; one JSON line per configuration: the per-round decode ms per token step (CUDA events on
the launch stream after one warm-up decode that also captures the graph), their minimum, and whether the greedy ids
equal those of the first configuration."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--max-len", type=int, default=300)
ap.add_argument("--policy", default="split24")
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--cfg", action="append", default=[], help="tag:name=value,name=value")
ap.add_argument("--defaults", default="", help="name=value,... applied before every configuration")
ap.add_argument("--out", default="")
args = ap.parse_args()


def parse(s):
    return [(kv.split("=")[0], int(kv.split("=")[1])) for kv in s.split(",") if kv]


cfgs = []
for c in args.cfg:
    tag, _, rest = c.partition(":")
    cfgs.append((tag, parse(rest)))
defaults = parse(args.defaults)
named = sorted({k for _, o in cfgs for k, _ in o})
if any(k not in dict(defaults) for k in named):
    missing = [k for k in named if k not in dict(defaults)]
    raise SystemExit(f"--defaults must give the reset value of every option used: missing {missing}")

B = args.batch
eng = Engine(synth.synthetic_state_dict(), device=0, max_batch=B, max_new_tokens=args.max_len, policy=args.policy)
wave = synth.synthetic_waveforms(2 * B).cuda(); ids = synth.synthetic_prompt_ids(B).cuda()
s = torch.cuda.Stream()
res = {tag: {"tag": tag, "opts": [f"{k}={v}" for k, v in o], "ms": []} for tag, o in cfgs}
ref = None
with torch.cuda.stream(s):
    eng.encode(wave[:B], wave[B:])
    for r in range(args.rounds):
        for tag, opts in cfgs:
            for k, v in defaults: eng.set_option(k, v)
            for k, v in opts: eng.set_option(k, v)
            eng.prefix(ids); eng.prefill(B, want_logits=False)
            toks = eng.decode(B, args.max_len)                       # warm-up + graph capture
            torch.cuda.synchronize()
            t = 0.0
            for _ in range(args.iters):
                eng.prefix(ids); eng.prefill(B, want_logits=False)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                toks = eng.decode(B, args.max_len)
                e1.record(s)
                torch.cuda.synchronize()
                t += e0.elapsed_time(e1)
            res[tag]["ms"].append(t / args.iters / toks.shape[1])
            tc = toks.cpu()
            if ref is None: ref = tc
            ne = ref != tc
            res[tag]["tokens_equal_first_config"] = bool(not ne.any())
            res[tag]["rows_differing"] = int(ne.any(dim=1).sum())
            res[tag]["first_differing_step"] = int(ne.any(dim=0).nonzero()[0]) if ne.any() else None
lines = []
for tag, _ in cfgs:
    d = res[tag]
    d.update(batch=B, steps=args.max_len, policy=args.policy, decode_ms_per_step_min=min(d["ms"]),
             decode_ms_per_step_mean=sum(d["ms"]) / len(d["ms"]))
    lines.append(json.dumps(d))
    print(lines[-1], flush=True)
if args.out:
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
eng.close()
