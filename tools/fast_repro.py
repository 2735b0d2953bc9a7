"""Reproducer for the round-1 open issue: policy fast + 32-column split-K decode tiles returned non-finite logits on the
first decode of a fresh handle in some processes (B=2, two handles in the process).  Run plain (repeat in a shell loop)
or under compute-sanitizer (--tool racecheck / initcheck / memcheck).  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sd = synth.synthetic_state_dict()
wave = synth.synthetic_waveforms(4); ids = synth.synthetic_prompt_ids(2)
e0 = Engine(sd, device=0, max_batch=4, max_new_tokens=32, policy="split")
e1 = Engine(None, device=0, max_batch=2, max_new_tokens=32, policy="fast", arena=e0.arena)
e1.set_option("wide_tiles", 1)
out = {}
for name, eng in (("split", e0), ("fast_wide", e1)):
    eng.encode(wave[:2], wave[2:]); eng.prefix(ids); eng.prefill(2, want_logits=False)
    toks, dump = eng.decode(2, steps, dump_logits=True)
    out[name] = {"finite": bool(torch.isfinite(dump).all()), "tokens": toks.cpu().tolist()[0][:4]}
    toks2 = eng.generate(wave[:2], wave[2:], ids, steps)
    out[name]["graph_tokens_equal"] = bool(torch.equal(toks2.cpu(), toks.cpu()))
print(json.dumps(out), flush=True)
