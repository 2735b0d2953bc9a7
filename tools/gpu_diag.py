"""First-light diagnostics on a GPU box: prints per-stage max errors vs the CPU oracle without stopping at the first
failure.  Usage: python tools/gpu_diag.py [policy]"""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine
from oracle import restated as R

policy = sys.argv[1] if len(sys.argv) > 1 else "split"
t0 = time.time()
sd = synth.synthetic_state_dict()
wave = synth.synthetic_waveforms(4); ids = synth.synthetic_prompt_ids(2)
w1, w2 = wave[:2], wave[2:]
golden = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "ref_synth1234.npz")))
eng = Engine(sd, device=0, max_batch=4, max_new_tokens=32, policy=policy)
print(f"engine up in {time.time()-t0:.1f}s, workspace {eng.workspace_bytes()/1e9:.2f} GB", flush=True)

def err(a, b):
    a = a.detach().cpu().double(); b = torch.as_tensor(np.asarray(b)).double()
    return (a - b).abs().max().item()

def step(name, fn):
    try:
        t = time.time(); r = fn(); torch.cuda.synchronize()
        print(f"[ok ] {name}: {r}  ({time.time()-t:.2f}s)", flush=True)
    except Exception as e:
        print(f"[ERR] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()

taps = {}
with torch.no_grad():
    ra = R.encode_clips(sd, w1, taps); rb = R.encode_clips(sd, w2)
    prefix = R.build_prefix(sd, ra, rb, ids)
    want_logits = R.last_logits(sd, R.llama_hidden(sd, prefix))

def gemm():
    out = []
    for (m, n, k) in [(300, 288, 96), (128, 960, 576), (5, 49152, 576), (77, 527, 4608)]:
        g = torch.Generator().manual_seed(1)
        a = torch.randn(m, k, generator=g); w = torch.randn(n, k, generator=g) / k ** 0.5
        out.append(round(err(eng.op_gemm(a, w), a.double() @ w.double().T), 7))
    return out
step("gemm max err", gemm)
def fe():
    lm, bn = eng.frontend(w1)
    return err(lm, taps["logmel"]), err(bn, taps["bn"])
step("frontend (logmel, bn)", fe)
for s, name in enumerate(["patch", "stage0", "stage1", "stage2", "stage3"]):
    step(f"tap {name}", lambda s=s, name=name: (err(eng.encode_tap(w1, s), taps[name]), float(taps[name].abs().max())))
def tail():
    lat, fr = eng.encode_tap(w1, 5)
    return err(lat, taps["latent"]), err(fr, taps["oframe"])
step("tail (latent, frames)", tail)
def rows():
    r = eng.encode(w1, w2)
    return err(r[0], ra), err(r[1], rb)
step("rows33", rows)
step("prefix", lambda: err(eng.prefix(ids), prefix))
def prefill():
    eng.set_prefix(prefix)
    lg = eng.prefill(2)
    return err(lg, want_logits), lg.argmax(-1).tolist(), golden["tokens"][:, 0].tolist()
step("prefill logits (oracle prefix)", prefill)
def decode_plain():
    os.environ["MB_NO_GRAPH"] = "1"
    eng.set_prefix(prefix); eng.prefill(2, want_logits=False)
    toks, dump = eng.decode(2, 12, dump_logits=True)
    probe = torch.from_numpy(golden["probe_ids"])
    e = err(dump.cpu()[:, :, probe], golden["probe_logits"])
    return e, toks.cpu().tolist() == golden["tokens"].tolist(), toks.cpu().tolist()
step("decode no-graph (logit err, ids match)", decode_plain)
def decode_graph():
    os.environ.pop("MB_NO_GRAPH", None)
    eng.set_prefix(prefix); eng.prefill(2, want_logits=False)
    toks = eng.decode(2, 12)
    return toks.cpu().tolist() == golden["tokens"].tolist(), toks.cpu().tolist()
step("decode graph (ids match)", decode_graph)
def e2e():
    toks = eng.generate(w1, w2, ids, 12)
    return toks.cpu().tolist() == golden["tokens"].tolist(), toks.cpu().tolist(), eng.kernel_launches
step("generate e2e", e2e)
print("golden tokens", golden["tokens"].tolist())
