"""Seeded synthetic checkpoint and inputs with the reference's exact state-dict schema.

There is no network and no real ``v0.ckpt`` / ``v0_s.ckpt`` on the build or GPU
boxes, so benches and parity tests run on a synthetic checkpoint that has the
same keys, shapes and dtypes the reference loads (reference
``mellow/wrapper.py:74-82``; SURVEY.md section 8a' and 8d).  Values come from
numpy's PCG64 integer stream mapped affinely to float32, which is bit-identical
on every machine (the golden fixtures under ``tests/golden`` were produced from
``synthetic_state_dict(seed=1234)`` and must be reproducible on the GPU box).
"""
import math

import numpy as np
import torch

from . import dsp
from . import schema as S

DEFAULT_SEED = 1234


class _Stream:
    def __init__(self, seed):
        self.rng = np.random.Generator(np.random.PCG64(seed))

    def uniform(self, shape, std):
        """float32 uniform with the given standard deviation (exactly reproducible)."""
        n = int(np.prod(shape)) if len(shape) else 1
        raw = self.rng.integers(0, 1 << 24, size=n, dtype=np.uint32).astype(np.float32)
        x = (raw - np.float32(8388608.0)) * np.float32(1.0 / 8388608.0)        # [-1, 1)
        x = x * np.float32(std * math.sqrt(3.0))
        return torch.from_numpy(x.reshape(shape))


def synthetic_state_dict(seed=DEFAULT_SEED):
    st = _Stream(seed)
    sd = {}
    real, imag = dsp.windowed_dft()
    relidx = torch.from_numpy(dsp.relative_position_index())
    for key, (shape, dtype) in S.checkpoint_schema().items():
        leaf = key.rsplit(".", 1)[-1]
        if key.endswith("conv_real.weight"):
            t = torch.from_numpy(real)[:, None, :].clone()
        elif key.endswith("conv_imag.weight"):
            t = torch.from_numpy(imag)[:, None, :].clone()
        elif key.endswith("melW"):
            t = torch.from_numpy(dsp.slaney_mel_matrix())
        elif key.endswith("relative_position_index"):
            t = relidx.clone()
        elif key.endswith("attn_mask"):
            res = int(round(math.sqrt(shape[0]))) * S.WINDOW
            t = torch.from_numpy(dsp.shifted_window_mask(res))
        elif key.endswith("num_batches_tracked"):
            t = torch.tensor(1000, dtype=torch.int64)
        elif key.endswith("bn0.running_mean"):
            t = -25.0 + st.uniform(shape, 4.0)
        elif key.endswith("bn0.running_var"):
            t = 60.0 + st.uniform(shape, 10.0)
        elif key.endswith("relative_position_bias_table"):
            t = st.uniform(shape, 0.5)
        elif key.endswith("lm_head.weight"):
            t = sd["caption_decoder.lm.model.embed_tokens.weight"]     # tied (same storage, like HF)
        elif key.endswith("embed_tokens.weight"):
            # 0.2 balances the token-embedding and layer contributions in the residual stream, so greedy decoding
            # wanders over the vocabulary (top-1/top-2 margins from ~1e-2 to ~2) instead of collapsing to one token
            t = st.uniform(shape, 0.2)
        elif key.endswith(("q_proj.weight", "k_proj.weight")):
            # gain 2 keeps attention content-dependent over the highly repetitive 389-token prefix (without it the
            # synthetic LM emits one token forever); larger gains make the random network chaotic
            t = st.uniform(shape, 2.0 / math.sqrt(shape[1]))
        elif leaf == "weight" and len(shape) == 1:                       # LayerNorm / RMSNorm / BN gain
            t = 1.0 + st.uniform(shape, 0.1)
        elif leaf == "bias":
            t = st.uniform(shape, 0.05)
        elif leaf == "weight":
            fan_in = int(np.prod(shape[1:]))
            t = st.uniform(shape, 1.0 / math.sqrt(fan_in))
        else:
            raise KeyError(key)
        assert tuple(t.shape) == tuple(shape), (key, t.shape, shape)
        sd[key] = t.to(getattr(torch, dtype))
    return sd


def synthetic_waveforms(n_clips, seed=DEFAULT_SEED):
    """(n_clips, 320000) float32: 0.1*uniform noise plus three slow chirped tones per clip, already 32 kHz / 10 s
    (no resample / tile / crop needed)."""
    st = _Stream(seed + 17)
    noise = st.uniform((n_clips, S.CLIP_SAMPLES), 0.1).numpy()
    t = np.arange(S.CLIP_SAMPLES, dtype=np.float64) / S.SAMPLE_RATE
    par = np.random.Generator(np.random.PCG64(seed + 18)).integers(0, 1 << 16, size=(n_clips, 3, 2))
    out = noise.astype(np.float64)
    for c in range(n_clips):
        for k in range(3):
            f0 = 100.0 + 6000.0 * par[c, k, 0] / 65536.0
            sweep = 400.0 * (par[c, k, 1] / 65536.0 - 0.5)
            out[c] += 0.2 * np.sin(2.0 * np.pi * (f0 * t + 0.5 * sweep * t * t / 10.0))
    return torch.from_numpy(np.clip(out, -1.0, 1.0).astype(np.float32))


def synthetic_prompt_ids(batch, n_real=64, pad_id=17, seed=DEFAULT_SEED):
    """(batch, 129) int64: ``n_real`` ids uniform in [18, VOCAB) then right padding with ``pad_id``
    (the id of '!' in the SmolLM2 vocabulary; reference pads with '!' at wrapper.py:85,186-190)."""
    rng = np.random.Generator(np.random.PCG64(seed + 29))
    ids = np.full((batch, S.TEXT_LEN), pad_id, dtype=np.int64)
    ids[:, :n_real] = rng.integers(18, S.VOCAB, size=(batch, n_real))
    return torch.from_numpy(ids)
