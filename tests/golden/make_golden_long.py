"""Generates the long-run golden fixtures by running the UNMODIFIED reference model classes (imported from
/root/reference through oracle/reference_model.py) under the reference's own cache-less decode loop
(mellow/wrapper.py:197-256, restated verbatim in oracle/reference_model.reference_generate_ids).  Build container only:

    python tests/golden/make_golden_long.py [set ...]         # default: every set

Sets (BASELINE.json configs):
  long1234   checkpoint seed 1234, the B=2 inputs of ref_synth1234.npz, 300 greedy steps (ctx 389..688: every key-tile
             boundary of the decode-attention kernel, the max_len of configs[2] / example.py:30)
  rows1234   checkpoint seed 1234, 4 further DISTINCT rows (input seed 4321) with ragged prompts (1, 5, 64 and 129 real
             tokens), 300 steps
  rows77     a second checkpoint (seed 77), 2 rows (input seed 99), 300 steps
  config0    configs[0]: resource/1.wav + resource/2.wav (copied to tests/golden/resource/), random.seed(0), the
             prompt of example.py:25, 30 greedy steps; audio prepared exactly like wrapper.py:141-168 (torchaudio
             Resample 44.1k -> 32k, tile / random crop), prompt tokenised with the offline stand-in tokenizer

Each fixture holds the token ids, the per-step top-8 (ids, values) and 64 probe logits of the reference.
"""
import os
import random
import shutil
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mellow_b200 import synth  # noqa: E402
from mellow_b200 import schema as S  # noqa: E402
from oracle.reference_model import REFERENCE_ROOT, build_reference_model, reference_generate_ids  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CONFIG0_PROMPT = ("what is the primary sound event present in the clip? a) dog barking b) chirping birds c) car engine "
                  "d) clapping")                                                    # example.py:25
PROBE = np.random.Generator(np.random.PCG64(7)).integers(0, S.VOCAB, size=64).astype(np.int64)


def ragged_prompt_ids(n_real_per_row, seed, pad_id=17):
    """(B,129) int64 with a different number of real ids per row (right padding like wrapper.py:186-190)."""
    rng = np.random.Generator(np.random.PCG64(seed + 29))
    ids = np.full((len(n_real_per_row), S.TEXT_LEN), pad_id, dtype=np.int64)
    for r, n in enumerate(n_real_per_row):
        ids[r, :n] = rng.integers(18, S.VOCAB, size=n)
    return torch.from_numpy(ids)


def set_inputs(name):
    """-> (checkpoint seed, wave1, wave2, ids, steps); shared with tests/test_gpu_long_parity.py."""
    if name == "long1234":
        w = synth.synthetic_waveforms(4)
        return 1234, w[:2], w[2:], synth.synthetic_prompt_ids(2), 300
    if name == "rows1234":
        w = synth.synthetic_waveforms(8, seed=4321)
        return 1234, w[:4], w[4:], ragged_prompt_ids([1, 5, 64, 129], seed=4321), 300
    if name == "rows77":
        w = synth.synthetic_waveforms(4, seed=99)
        return 77, w[:2], w[2:], ragged_prompt_ids([20, 100], seed=99), 300
    raise KeyError(name)


def run_reference(model, wave1, wave2, ids, steps):
    with torch.no_grad():
        prefix, _, _ = model.generate_prefix_inference({"audio1": wave1, "audio2": wave2, "input": {"input_ids": ids}})
        toks, logits = reference_generate_ids(model, prefix, steps, top_p=0.8, temperature=1.0, dump_logits=True)
    logits = torch.stack(logits, 0)                                    # (steps, B, V)
    top = logits.topk(8, dim=-1)
    return {"tokens": toks.numpy().astype(np.int64), "top8_ids": top.indices.numpy().astype(np.int64),
            "top8_vals": top.values.numpy(), "probe_ids": PROBE,
            "probe_logits": logits[:, :, torch.from_numpy(PROBE)].numpy(), "input_ids": ids.numpy().astype(np.int64)}


def make_synth(name):
    seed, w1, w2, ids, steps = set_inputs(name)
    model = build_reference_model(synth.synthetic_state_dict(seed))
    return run_reference(model, w1, w2, ids, steps)


def make_config0():
    from mellow_b200.audio_io import load_audio_into_tensor
    from mellow_b200.tokenizer import ByteStandInTokenizer, tokenize_prompts
    res = os.path.join(HERE, "resource")
    os.makedirs(res, exist_ok=True)
    for f in ("1.wav", "2.wav"):                                       # input fixtures of the reference (resource/*.wav)
        if not os.path.isfile(os.path.join(res, f)):
            shutil.copyfile(os.path.join(REFERENCE_ROOT, "resource", f), os.path.join(res, f))
    random.seed(0)                                                      # draws: list 1 first, then list 2 (wrapper.py:277-278)
    a1 = load_audio_into_tensor(os.path.join(res, "1.wav"), 10, 32000, True, random)[None]
    a2 = load_audio_into_tensor(os.path.join(res, "2.wav"), 10, 32000, True, random)[None]
    ids = tokenize_prompts(ByteStandInTokenizer(), [CONFIG0_PROMPT], S.TEXT_LEN)
    model = build_reference_model(synth.synthetic_state_dict())
    g = run_reference(model, a1, a2, ids, 30)
    g["audio1_head"] = a1[0, :4096].numpy()
    g["audio2_head"] = a2[0, :4096].numpy()
    return g


def main():
    torch.manual_seed(0)
    names = sys.argv[1:] or ["config0", "long1234", "rows77", "rows1234"]
    for name in names:
        t0 = time.time()
        g = make_config0() if name == "config0" else make_synth(name)
        out = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(out, **g)
        margins = g["top8_vals"][..., 0] - g["top8_vals"][..., 1]
        print(f"{name}: wrote {out} ({os.path.getsize(out)} bytes) in {time.time() - t0:.0f} s; tokens[:, :8] = "
              f"{g['tokens'][:, :8].tolist()}; min top1-top2 margin {margins.min():.5f} at step "
              f"{int(np.unravel_index(margins.argmin(), margins.shape)[0])}; margins < 0.02: {(margins < 0.02).sum()}", flush=True)


if __name__ == "__main__":
    main()
