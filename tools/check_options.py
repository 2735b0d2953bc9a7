"""Debug aid: prefill logits of one option setting against the CPU oracle (B=2), one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine
from oracle import restated as R
opts = [kv.split("=") for kv in sys.argv[1:]]
sd = synth.synthetic_state_dict()
g = torch.Generator().manual_seed(5)
prefix = torch.randn(2, 389, 576, generator=g) * 0.3
with torch.no_grad():
    want = R.last_logits(sd, R.llama_hidden(sd, prefix))
eng = Engine(sd, device=0, max_batch=4, max_new_tokens=16, policy="split24")
for k, v in opts:
    eng.set_option(k, int(v))
eng.set_prefix(prefix)
got = eng.prefill(2).cpu()
torch.cuda.synchronize()
out = {"opts": dict(opts), "prefill_max_err": (got - want).abs().max().item(), "finite": bool(torch.isfinite(got).all())}
toks = eng.decode(2, 8).cpu()
out["decode_tokens_row0"] = toks[0].tolist()
print(json.dumps(out), flush=True)
