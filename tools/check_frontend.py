"""Small encoder pass for compute-sanitizer: one pair of synthetic clips through mb_encode (log-mel front end, patch embed,
Swin stages, tail), prints whether the rows are finite.  tools/gpu_sanitize.sh runs it under memcheck / racecheck."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine
eng = Engine(synth.synthetic_state_dict(), device=0, max_batch=2, max_new_tokens=8, policy="split24")
wave = synth.synthetic_waveforms(2).cuda()
rows = eng.encode(wave[:1], wave[1:])
torch.cuda.synchronize()
print(json.dumps({"rows_finite": bool(torch.isfinite(rows).all()), "rows_abs_mean": float(rows.abs().mean())}))
eng.close()
