#!/usr/bin/env python
"""Benchmark of the Mellow two-audio-plus-prompt inference path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle) on the host cores

    python bench.py --workload configs3|configs4|configs1 ...  # the other BASELINE.json configurations (see WORKLOADS)

Default workload (BASELINE.json configs[2], the batch-128 configuration the metric is quoted on): per GPU, B = 128 synthetic
pairs of 10 s / 32 kHz clips, 64-token prompts padded to 129, max_len = 300, top_p = 0.8, temperature = 1.0, seeded
synthetic checkpoint with the reference's schema (no real weights offline).  One "step" = one full generate() over the
batch: log-mel front end, HTSAT over 2*B clips, projection/prefix, 389-token LM prefill, 300 KV-cached decode steps.

value  = generated tokens / s over the whole job with inputs resident in HBM (CUDA events on the launch stream).
e2e    = the same through MellowWrapper's engine call with HOST (pinned) inputs, H2D + D2H inside the timed region.
phases = prefill pairs/s (front end + encoder + prefix + LM prefill) and decode tokens/s, timed separately.
roofline = the decode-attention kernel (the dominant kernel of the decode loop): algorithmic KV bytes per launch /
           its CUDA-event time, against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline = the oracle (reference algorithm restated, cache-less loop) on a bounded sample, rank 0, N = 1 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

PREFIX, LAYERS, KV_HEADS, HEAD_DIM, HIDDEN = 389, 30, 3, 64, 576
FALLBACK_HBM_GBS = 6650.0


FALLBACK_BF16_TFLOPS = 1400.0          # sustained figure of the profiling recipe (1.59 PF/s burst)
LM_PARAMS = 134_515_008


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        if "bf16_tflops_sustained" in d:
            return float(d["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return FALLBACK_BF16_TFLOPS, "fallback (B200_PROFILING.md ~1.4 PF/s sustained)"


# BASELINE.json `configs` by index.  "weak": every GPU gets `batch` pairs; "strong": `batch` pairs in total, split over
# the GPUs (contiguous slices, like MellowWrapper.generate under torchrun).
WORKLOADS = {
    "configs2": {"batch": 128, "max_len": 300, "scaling": "weak",
                 "name": "BASELINE.json configs[2]: v0_s batch-128 two-audio difference, max_len=300, top_p=0.8, temp=1.0"},
    "configs1": {"batch": 32, "max_len": 30, "scaling": "weak",
                 "name": "BASELINE.json configs[1]: v0 batch-32 captioning prompts, 10 s clips, greedy decode (max_len=30)"},
    "configs3": {"batch": 512, "max_len": 8, "scaling": "strong",
                 "name": "BASELINE.json configs[3]: v0 batch-512 MCQ (prefill-heavy, 8-token decode), batch-sharded"},
    "configs4": {"batch": 256, "max_len": 300, "scaling": "strong",
                 "name": "BASELINE.json configs[4]: v0_s batch-256 mixed ReasonAQA prompts, max_len=300, batch-sharded"},
}


def decode_floor_bytes(batch, max_len, elem_bytes, kv_bytes=None):
    """bytes one decode step must move, summed over the steps 1..max_len-1 that run the LM
    (step 0 samples from the prefill logits): every LM weight once, the KV history read, the new KV row written, the
    token embedding gathered; `elem_bytes` = bytes per weight / KV element of the policy (4 under `split`)."""
    kv_bytes = elem_bytes if kv_bytes is None else kv_bytes
    total = 0
    for t in range(1, max_len):
        ctx = PREFIX + t
        total += elem_bytes * (LM_PARAMS + HIDDEN * batch) + kv_bytes * (11520 * batch * ctx + 11520 * batch)
    return total


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        # samples under load: upper half of the distribution (idle samples before/after the region are lower)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_sample(pairs, steps, threads):
    """Reference algorithm on the host: encoder x2 + prefix + `steps` iterations of the cache-less loop."""
    from mellow_b200 import synth
    from oracle import restated as R        # the only place bench.py touches oracle/: the CPU baseline legs
    torch.set_num_threads(threads)
    sd = synth.synthetic_state_dict()
    wave = synth.synthetic_waveforms(2 * pairs)
    ids = synth.synthetic_prompt_ids(pairs)
    with torch.no_grad():
        R.encode_clips(sd, wave[:1])                                  # warm the thread pool / allocator
        t0 = time.perf_counter()
        prefix = R.build_prefix(sd, R.encode_clips(sd, wave[:pairs]), R.encode_clips(sd, wave[pairs:]), ids)
        t1 = time.perf_counter()
        toks = R.generate_ids(sd, prefix, steps, top_p=0.8, temperature=1.0)
        t2 = time.perf_counter()
    n_tok = toks.numel()
    return {"tokens": n_tok, "total_s": t2 - t0, "prefix_s": t1 - t0, "loop_s": t2 - t1,
            "tokens_per_s": n_tok / (t2 - t0), "decode_tokens_per_s": n_tok / (t2 - t1),
            "prefix_pairs_per_s": pairs / (t1 - t0)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    pairs, steps = 2, 6
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_oracle_sample(pairs, steps, cores)
        if i >= args.warmup:
            vals.append(r)
    tps = sum(v["tokens"] for v in vals) / sum(v["total_s"] for v in vals)
    ms = 1e3 * sum(v["total_s"] for v in vals) / len(vals)
    sample = (f"B={pairs} pairs, encoder x2 + prefix + {steps} steps of the reference's cache-less loop (ctx 389..{388 + steps}), "
              "fp32, torch CPU; the full batch-128 x 300-step reference run is ~6.4 PFLOP and is not runnable on the host")
    line = {"impl": "reference", "metric": "generated tokens/s, generate() = prefill + decode", "value": tps, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["name"], "sampled_as": sample},
            "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample,
                             "decode_tokens_per_s": sum(v["tokens"] for v in vals) / sum(v["loop_s"] for v in vals),
                             "prefill_pairs_per_s": pairs * len(vals) / sum(v["prefix_s"] for v in vals)},
            "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    from mellow_b200 import synth
    from mellow_b200.dist import build_engine, env_world, shard_bounds
    wl = WORKLOADS[args.workload]
    max_len = args.max_len or wl["max_len"]
    total_batch = args.batch or wl["batch"]
    rank0, _, world0 = env_world()
    if wl["scaling"] == "strong":                       # a fixed job split over the ranks (contiguous slices)
        lo, hi = shard_bounds(total_batch, rank0, world0)
        B = hi - lo
    else:                                               # every rank gets its own `batch` pairs
        lo, B = 0, total_batch
    cap = min(B, 128)                                   # pairs per engine pass (one 128-row tile of the decode GEMMs)
    # NCCL announces its version on stdout at communicator creation; the contract is ONE JSON line on stdout, so file
    # descriptor 1 points at stderr while the process group and the weight broadcast are set up
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        eng, rank, world = build_engine(synth.synthetic_state_dict, max_batch=cap, max_new_tokens=max(max_len, 8), policy=args.policy)
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    dev = eng.device
    torch.cuda.set_device(dev)
    # weak: a different seed per rank (every rank processes its own pairs); strong: one seeded job, this rank's slice
    seed = 1234 + (rank if wl["scaling"] == "weak" else 0)
    n_all = total_batch if wl["scaling"] == "strong" else B
    wave = synth.synthetic_waveforms(2 * n_all, seed=seed)
    ids = synth.synthetic_prompt_ids(n_all, seed=seed)
    w1_h, w2_h = wave[lo:lo + B].contiguous().pin_memory(), wave[n_all + lo:n_all + lo + B].contiguous().pin_memory()
    ids_h = ids[lo:lo + B].to(torch.int32).pin_memory()
    out_h = torch.empty(B, max_len, dtype=torch.int32).pin_memory()
    w1_d, w2_d, ids_d = w1_h.to(dev), w2_h.to(dev), ids_h.to(dev)
    stream = torch.cuda.Stream(device=dev)
    spans = [(s, min(B, s + cap)) for s in range(0, B, cap)]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, iters, warmup):
        """fn() launched on `stream`; returns (ms per iteration as max over ranks, last result)."""
        res = None
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                res = fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(iters):
                res = fn()
            e1.record(stream)
            barrier()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, res

    def gen_device():
        return sum(eng.generate(w1_d[a:b], w2_d[a:b], ids_d[a:b], max_len, temperature=1.0, top_p=0.8).shape[1] * (b - a)
                   for a, b in spans)

    def gen_host():
        return sum(eng.generate_host(w1_h[a:b], w2_h[a:b], ids_h[a:b], max_len, temperature=1.0, top_p=0.8,
                                     out=out_h[a:b]).shape[1] * (b - a) for a, b in spans)

    def total_over_ranks(n_local):
        if world == 1:
            return n_local
        t = torch.tensor([float(n_local)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    launches0 = eng.kernel_launches
    sampler = ClockSampler(dev.index)
    sampler.start()
    ms_dev, tokens_local = timed(gen_device, args.steps, args.warmup)
    clocks = sampler.stop()
    launches_per_step = (eng.kernel_launches - launches0) // (args.steps + args.warmup)
    tokens_per_step = total_over_ranks(tokens_local)              # generated tokens of the whole job per step
    value = tokens_per_step / (ms_dev * 1e-3)

    ms_e2e, tokens_local_h = timed(gen_host, args.steps, args.warmup)
    e2e_value = total_over_ranks(tokens_local_h) / (ms_e2e * 1e-3)

    # phases (not part of `value`): this rank's first pass
    a0, b0 = spans[0]
    Bp = b0 - a0
    # straight C-ABI calls without output buffers: nothing is allocated inside the timed regions; 2 warm-ups + 4 iterations
    # (the first tensor-heavy kernels after the memory-bound decode loop see a transient power-management clock dip)
    import ctypes
    st_ptr = ctypes.c_void_p(stream.cuda_stream)
    pv = lambda t: ctypes.c_void_p(t.data_ptr())
    w1p, w2p, idp = w1_d[a0:b0].contiguous(), w2_d[a0:b0].contiguous(), ids_d[a0:b0].contiguous()

    def enc_only():
        eng._ck(eng.lib.mb_encode(eng.handle, pv(w1p), pv(w2p), Bp, None, st_ptr))
        eng._ck(eng.lib.mb_prefix(eng.handle, pv(idp), Bp, None, st_ptr))
    # encoder and LM prefill are timed in sequence, as generate() runs them (CUDA events between the two phases on the
    # launch stream): back-to-back repeats of the LM prefill alone sit at the sustained power-capped tensor clock and read
    # ~15 % slower than the same kernels do inside a generate()
    lm_only = lambda: eng._ck(eng.lib.mb_prefill(eng.handle, Bp, None, st_ptr))
    for _ in range(2):
        enc_only(); lm_only()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
    with torch.cuda.stream(stream):
        evs[0].record(stream)
        for i in range(4):
            enc_only(); evs[2 * i + 1].record(stream)
            lm_only(); evs[2 * i + 2].record(stream)
    torch.cuda.synchronize()
    ms_enc = sum(evs[2 * i].elapsed_time(evs[2 * i + 1]) for i in range(4)) / 4
    ms_lm = sum(evs[2 * i + 1].elapsed_time(evs[2 * i + 2]) for i in range(4)) / 4
    ms_prefill = ms_enc + ms_lm
    ms_decode, dtoks = timed(lambda: eng.decode(Bp, max_len), 1, 1)
    # prefill is tensor-pipe bound: 112.18 GFLOP of algorithmic work per pair = 2 x 12.14 (encoder) + 87.90 (LM), SURVEY.md
    # section 8d; under the split policy every contraction is issued as 3 bf16 MMA passes
    split = args.policy in ("split", "split24")
    passes = 3 if split else 1
    tpeak, tpeak_src = measured_tensor_peak()
    hpeak, hpeak_src = measured_peaks()
    kv_bytes = 4 if args.policy == "split" else (3 if args.policy == "split24" else 2)
    w_bytes = 4 if split else 2
    n_dec = max(max_len - 1, 1)
    floor_own = decode_floor_bytes(Bp, max_len, w_bytes, kv_bytes) / (hpeak * 1e9) * 1e3     # what THIS policy must move
    floor_4b = decode_floor_bytes(Bp, max_len, 4) / (hpeak * 1e9) * 1e3                      # fp32 weights and KV
    floor_2b = decode_floor_bytes(Bp, max_len, 2) / (hpeak * 1e9) * 1e3
    alg_tf = lambda gf, ms: gf * 1e9 * Bp / (ms * 1e-3) / 1e12
    phases = {"pairs": Bp, "prefill_ms": ms_prefill, "encoder_ms": ms_enc, "lm_prefill_ms": ms_lm,
              "prefill_pairs_per_s": Bp / (ms_prefill * 1e-3),
              "prefill_algorithmic_tflops_per_gpu": alg_tf(112.18, ms_prefill),
              "encoder_algorithmic_tflops_per_gpu": alg_tf(24.28, ms_enc),
              "lm_prefill_algorithmic_tflops_per_gpu": alg_tf(87.90, ms_lm),
              "prefill_frac_of_tensor_peak_algorithmic": alg_tf(112.18, ms_prefill) / tpeak,
              "prefill_frac_of_tensor_peak_issued_mma": alg_tf(112.18, ms_prefill) * passes / tpeak,
              "mma_passes": passes, "tensor_peak_tflops": tpeak, "tensor_peak_source": tpeak_src,
              "decode_ms_per_token_step": ms_decode / max_len,
              "decode_tokens_per_s": Bp * dtoks.shape[1] / (ms_decode * 1e-3),
              "decode_frac_of_hbm_floor": {"policy_bytes": floor_own / ms_decode, "fp32_4B": floor_4b / ms_decode,
                                           "bf16_2B_BASELINE_md": floor_2b / ms_decode},
              "decode_hbm_floor_ms_per_step": {"policy_bytes": floor_own / n_dec, "fp32_4B": floor_4b / n_dec,
                                               "bf16_2B_BASELINE_md": floor_2b / n_dec},
              "decode_floor_note": f"SURVEY 8d bytes summed over the loop; policy {args.policy}: {w_bytes} B per weight, "
                                   f"{kv_bytes} B per KV element; peak {hpeak:.0f} GB/s ({hpeak_src})"}

    # roofline of the dominant decode kernel: decode attention at the mean context of the loop
    peak, peak_src = measured_peaks()
    ctx = PREFIX + max_len // 2
    iters = 120
    ms_attn, _ = timed(lambda: eng.bench_decode_attention(Bp, ctx, iters), 1, 1)
    ms_attn /= iters
    alg_bytes = 2 * Bp * KV_HEADS * ctx * HEAD_DIM * kv_bytes + 2 * Bp * HIDDEN * 4
    achieved = alg_bytes / (ms_attn * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tname = {3: "r2_decode_attention_final_tree_b128_ctx539_ncu_full.json"}.get(kv_bytes)
    tpath = os.path.join(ROOT, "profiles", tname) if tname else ""
    if tpath and os.path.isfile(tpath) and Bp == 128 and ctx == 539:
        with open(tpath) as f:                       # dram__bytes_read+write per launch from the committed ncu --set full capture
            traffic = json.load(f)["traffic_bytes_per_launch"]
        traffic_src = f"cited from profiles/{tname} (ncu --set full of this kernel at this operating point), not measured in this run"
    roofline = {"kernel": "decode_attention_warp_kernel<kv24>" if kv_bytes == 3 else "decode_attention_warp_kernel", "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "us_per_launch": ms_attn * 1e3, "ctx": ctx,
                "timing": "CUDA events over 120 back-to-back launches cycling the 30 layer caches (successive launches "
                          "overlap through PDL; an isolated ncu launch is slower, see profiles/)",
                "share_of_decode_step": LAYERS * ms_attn / (ms_decode / max_len)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        r = cpu_oracle_sample(2, 6, cores)
        cpu = {"value": r["tokens_per_s"], "unit": "tokens/s", "cores": cores, "kind": "port",
               "sample": "B=2 pairs, encoder x2 + prefix + 6 steps of the reference's cache-less loop (ctx 389..394), fp32 torch CPU",
               "decode_tokens_per_s": r["decode_tokens_per_s"], "prefill_pairs_per_s": r["prefix_pairs_per_s"]}

    if rank == 0:
        pairs_all = total_batch if wl["scaling"] == "strong" else B * world
        line = {"metric": "generated tokens/s, generate() = prefill + decode", "value": value, "unit": "tokens/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
                "scaling": wl["scaling"], "vs_baseline": None,
                "dtype": {"split": "bf16x3 (bf16 hi/lo split operands, fp32 accumulate, fp32 KV)",
                          "split24": "bf16x3 (bf16 hi/lo split operands, fp32 accumulate, 24-bit KV)", "fast": "bf16"}[args.policy],
                "data": "synthetic",
                "config": {"workload": wl["name"], "pairs_total": pairs_all, "pairs_per_gpu": B, "max_len": max_len,
                           "tokens_per_step": tokens_per_step, "prompt_tokens": 64, "policy": args.policy,
                           "passes_per_gpu": len(spans), "pairs_per_s": pairs_all / (ms_dev * 1e-3),
                           "checkpoint": "synthetic seed 1234 (reference schema)",
                           "l2": "per-step inputs (2 x 164 MB waveforms) and the KV stream (>2 GB/step) exceed the 126 MB L2",
                           "parallelism": f"batch-sharded replicas x{world}, one NCCL weight broadcast at init, no per-step collective"},
                "e2e": {"value": e2e_value, "unit": "tokens/s", "ms_per_step": ms_e2e,
                        "h2d_bytes_per_step": int(2 * B * 320000 * 4 + B * 129 * 4), "d2h_bytes_per_step": int(B * max_len * 4)},
                "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_per_step": int(launches_per_step),
                "clocks": clocks, "phases": phases, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="configs2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the workload's pairs (per GPU if weak, total if strong)")
    ap.add_argument("--max-len", type=int, default=0, help="override the workload's max_len")
    ap.add_argument("--policy", default="split24", choices=["split", "split24", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU oracle")
        run_ours(args)


if __name__ == "__main__":
    main()
