"""Replicates tests/test_gpu_parity.py::test_fast_policy_logits_within_bf16_tolerance with diagnostics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine
from oracle import restated as R
sd = synth.synthetic_state_dict()
wave = synth.synthetic_waveforms(4); ids = synth.synthetic_prompt_ids(2)
with torch.no_grad():
    prefix = R.build_prefix(sd, R.encode_clips(sd, wave[:2]), R.encode_clips(sd, wave[2:]), ids)
eng = Engine(sd, device=0, max_batch=4, max_new_tokens=32, policy="split")
fast = Engine(None, device=0, max_batch=2, max_new_tokens=32, policy="fast", arena=eng.arena)
for trial in range(2):
    fast.set_prefix(prefix)
    logits = fast.prefill(2)
    toks, dump = fast.decode(2, 6, dump_logits=True)
    torch.cuda.synchronize()
    fin = torch.isfinite(dump)
    print("trial", trial, "prefill finite", bool(torch.isfinite(logits).all()), "tokens", toks.cpu().tolist())
    for s in range(dump.shape[0]):
        for b in range(2):
            n_bad = int((~fin[s, b]).sum())
            print(f"  step {s} row {b}: non-finite {n_bad} of {dump.shape[2]}  max|logit| {float(dump[s, b][fin[s, b]].abs().max()) if n_bad < dump.shape[2] else float('nan'):.3f}")
