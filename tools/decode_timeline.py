"""In-kernel timeline of the decode step (mb_set_trace): where the time of the ~210 dependent kernels goes.

    python tools/decode_timeline.py --batch 128 --steps 24 [--out gpurun_out/timeline.txt]

Every decode kernel stamps %globaltimer at entry of its first CTA, when that CTA's programmatic-dependency wait
returns, when it exits, and when the last CTA of the grid exits.  The tool prints one layer of one steady-state step
kernel by kernel and the per-kind averages over all layers of that step:
    launch->wait : entry of the first CTA to the return of griddepcontrol.wait (prologue overlapped with predecessor)
    dep gap      : predecessor's last-CTA exit to this kernel's wait return (what a kernel boundary costs)
    body         : wait return to last-CTA exit
"""
import argparse, os, struct, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import numpy as np
import torch
from mellow_b200 import synth
from mellow_b200.engine import Engine

KIND = {1: "qkv", 2: "attn", 3: "o_proj", 4: "add+norm", 5: "gate/up", 6: "down", 7: "add+norm2", 8: "lm_head"}
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--show-step", type=int, default=12)
ap.add_argument("--show-layer", type=int, default=15)
ap.add_argument("--out", default="")
ap.add_argument("--policy", default="split24")
ap.add_argument("--opt", action="append", default=[], help="name=value for mb_set_option")
args = ap.parse_args()
B = args.batch
eng = Engine(synth.synthetic_state_dict(), device=0, max_batch=B, max_new_tokens=max(args.steps, 8), policy=args.policy)
for kv in args.opt:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
wave = synth.synthetic_waveforms(2 * B).cuda(); ids = synth.synthetic_prompt_ids(B).cuda()
cap = args.steps * 230 * 2 * 2             # records of 16 events (first and last CTA of every launch)
buf = torch.zeros(8 + 256 * cap, dtype=torch.uint8, device="cuda")
lines = []
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    eng.encode(wave[:B], wave[B:]); eng.prefix(ids); eng.prefill(B, want_logits=False)
    eng.decode(B, args.steps)
    torch.cuda.synchronize()
    buf[:8] = torch.frombuffer(bytearray(struct.pack("II", 0, cap)), dtype=torch.uint8).cuda()
    eng._ck(eng.lib.mb_set_trace(eng.handle, ctypes.c_void_p(buf.data_ptr())))
    eng.encode(wave[:B], wave[B:]); eng.prefix(ids); eng.prefill(B, want_logits=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.decode(B, args.steps)              # captures the traced graph
    eng.encode(wave[:B], wave[B:]); eng.prefix(ids); eng.prefill(B, want_logits=False)
    torch.cuda.synchronize()
    buf.zero_()
    buf[:8] = torch.frombuffer(bytearray(struct.pack("II", 0, cap)), dtype=torch.uint8).cuda()
    e0.record(s)
    eng.decode(B, args.steps)
    e1.record(s)
    torch.cuda.synchronize()
    eng._ck(eng.lib.mb_set_trace(eng.handle, ctypes.c_void_p(0)))
raw = buf.cpu().numpy()
n = int(np.frombuffer(raw[:4].tobytes(), dtype=np.uint32)[0])
n = min(n, cap)
ev = np.frombuffer(raw[8:8 + 256 * n].tobytes(), dtype=np.dtype([("t", "<u8"), ("id", "<u4"), ("sm", "<u4")])).reshape(n, 16)
lines.append(f"records {n} (cap {cap}); traced decode {e0.elapsed_time(e1) / (args.steps - 1):.4f} ms/step over {args.steps - 1} graph steps")
t0 = int(ev["t"][ev["t"] > 0].min())
firsts, lasts = {}, {}
for r in range(n):
    ids = ev["id"][r]
    present = np.nonzero(ids)[0]
    if len(present) == 0:
        continue
    kid = int(ids[present[0]]) >> 4
    d = {"id": kid}
    for ph in present:
        d[int(ph)] = int(ev["t"][r][ph]) - t0
    if 0 in d:
        d["entry"] = d[0]
        firsts.setdefault(kid, []).append(d)
    else:
        lasts.setdefault(kid, []).append(d)
rows = []
for kid, lst in firsts.items():
    lst.sort(key=lambda d: d["entry"])
    ll = sorted(lasts.get(kid, []), key=lambda d: d.get(15, 0))
    for i, d in enumerate(lst):
        if 3 not in d and i < len(ll) and 3 in ll[i]:
            d[3] = ll[i][3]                    # last-CTA exit comes from the last CTA's own record
        rows.append(d)
rows.sort(key=lambda d: d["entry"])
# steps = runs between lm_head launches
steps, cur = [], []
for r in rows:
    cur.append(r)
    if r["id"] // 1000 == 8:
        steps.append(cur); cur = []
lines.append(f"steps seen {len(steps)}")
if len(steps) > args.show_step:
    st = steps[args.show_step]
    lines.append(f"step {args.show_step}: {len(st)} launches, {(st[-1].get(3, st[-1]['entry']) - st[0]['entry']) / 1e3:.1f} us first entry -> lm_head last exit")
    prev_end = None
    agg = {}
    lines.append(f"{'kernel':>14} {'entry':>9} {'waited':>9} {'exit0':>9} {'exitL':>9} | {'launch->wait':>12} {'dep gap':>8} {'body':>8}   (us, relative to step start)")
    base = st[0]["entry"]
    for r in st:
        kind, layer = r["id"] // 1000, r["id"] % 100
        w, x0, xl = r.get(1), r.get(2), r.get(3)
        end = max(v for v in (x0, xl) if v is not None) if (x0 is not None or xl is not None) else None
        lw = (w - r["entry"]) / 1e3 if w is not None else float("nan")
        gap = (w - prev_end) / 1e3 if (w is not None and prev_end is not None) else float("nan")
        body = (end - w) / 1e3 if (w is not None and end is not None) else float("nan")
        a = agg.setdefault(KIND.get(kind, str(kind)), [[], [], []])
        a[0].append(lw); a[1].append(gap); a[2].append(body)
        if layer in (args.show_layer, args.show_layer + 1) or kind == 8:
            f = lambda v: f"{(v - base) / 1e3:9.2f}" if v is not None else "        -"
            lines.append(f"{KIND.get(kind, kind):>11}.{layer:<2} {f(r['entry'])} {f(w)} {f(x0)} {f(xl)} | {lw:12.2f} {gap:8.2f} {body:8.2f}")
        if end is not None:
            prev_end = end
    # finer stamps of the weight-resident GEMM (first CTA): 4 weights of k-block 0 present, 5 first activation stage
    # landed, 6 all MMAs issued, 7 accumulator complete (seen by an epilogue warp), 8 that warp's stores issued
    fine = {}
    for r in st:
        if 1 in r and 5 in r and r["id"] // 1000 != 2:
            k = KIND.get(r["id"] // 1000, "?")
            f = fine.setdefault(k, [])
            f.append([(r.get(p, np.nan) - r[1]) / 1e3 for p in (4, 5, 6, 7, 8, 2, 3)])
    att = [[(r.get(p, np.nan) - r[1]) / 1e3 for p in (4, 5, 2, 3)] for r in st if r["id"] // 1000 == 2 and 1 in r]
    if att:
        lines.append("attention, first CTA, us after its wait returned: q/k/v ready (fused QKV mode), key tiles consumed, exit, last-CTA exit")
        lines.append("            " + " ".join(f"{x:7.2f}" for x in np.nanmean(np.array(att, dtype=float), axis=0)))
    if fine:
        lines.append("weight-resident GEMM, first CTA, us after its wait returned: W0 present, A0 landed, MMAs issued, acc complete, stores issued, exit, last-CTA exit")
        for k, v in fine.items():
            lines.append(f"{k:>12}: " + " ".join(f"{x:7.2f}" for x in np.nanmean(np.array(v), axis=0)))
    lines.append("per-kind means over the step (us): launch->wait, dep gap, body, n")
    tot = 0.0
    for k, (a, b, c) in agg.items():
        m = lambda v: float(np.nanmean(v)) if len(v) else float("nan")
        lines.append(f"{k:>12}: {m(a):7.2f} {m(b):7.2f} {m(c):7.2f}  n={len(a)}")
        tot += np.nansum(b) + np.nansum(c)
    lines.append(f"sum of (dep gap + body) over the step: {tot:.1f} us")
text = "\n".join(lines)
print(text)
if args.out:
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text + "\n")
    np.save(args.out + ".events.npy", ev)
eng.close()
