"""state_dict (reference schema, SURVEY.md section 8a') -> one contiguous weight arena laid out by the library.

The arena layout (entry names, offsets, sizes) is owned by ``libmellow_b200.so`` (``mb_weight_entry_*``); this
module produces the bytes of every entry:
  * GEMM weights as two bf16 planes ``hi = bf16(w)``, ``lo = bf16(w - hi)`` in nn.Linear ``[N, K]`` layout;
  * fused / re-ordered matrices: Swin qkv with the q rows pre-scaled by head_dim^-0.5 (reference htsat.py:311),
    SmolLM2 q|k|v stacked with the q/k head dims pair-interleaved (rotate-half partners adjacent), gate/up rows
    interleaved, the (2,3) TSCAM conv flattened to a [527, 4608] matrix, c2l K-padded 527 -> 544;
  * fp32 side tables: norm gains/biases, the BatchNorm folded to scale/shift, the 64x64 relative-position bias per
    head (table gathered through the checkpoint's own index buffer), RoPE cos/sin for 8192 positions, FFT twiddles.
It also verifies the assumptions the kernels bake in (``check_*``) and raises if a checkpoint violates them.
"""
import math

import numpy as np
import torch

from . import dsp
from . import schema as S

HT = "audio_encoder.base.htsat."
LMK = "caption_decoder.lm."
MAX_POS = 8192                     # rope table rows of the library (common.cuh kMaxPos)


def strip_module_prefix(sd):
    """The reference retries a strict load after dropping a DataParallel 'module.' prefix (wrapper.py:77-82)."""
    if all(k.startswith("module.") for k in sd):
        return {k[7:]: v for k, v in sd.items()}
    return sd


def check_schema(sd):
    want = S.checkpoint_schema()
    missing = [k for k in want if k not in sd]
    extra = [k for k in sd if k not in want]
    if missing or extra:
        raise RuntimeError(f"checkpoint does not match the Mellow schema: missing {missing[:4]} unexpected {extra[:4]}")
    for k, (shape, dtype) in want.items():
        if tuple(sd[k].shape) != tuple(shape):
            raise RuntimeError(f"{k}: shape {tuple(sd[k].shape)} != {tuple(shape)}")


def check_frontend_basis(sd, tol=2e-6):
    """The FFT front end is valid iff conv_real/conv_imag == window[n] * (cos, -sin)(2 pi n k / 1024)."""
    real = sd[HT + "spectrogram_extractor.stft.conv_real.weight"][:, 0, :].double().numpy()
    imag = sd[HT + "spectrogram_extractor.stft.conv_imag.weight"][:, 0, :].double().numpy()
    win = real[0]                                   # k = 0 row: cos = 1
    n = np.arange(S.N_FFT)[None, :]
    k = np.arange(S.N_BINS)[:, None]
    ph = 2.0 * np.pi * ((n * k) % S.N_FFT) / S.N_FFT
    err = max(np.abs(real - win * np.cos(ph)).max(), np.abs(imag + win * np.sin(ph)).max())
    if err > tol:
        raise RuntimeError(f"checkpoint STFT basis is not a windowed DFT (max deviation {err:.3g}); "
                           "the FFT front end does not apply to this checkpoint")
    return err


def check_shift_masks(sd):
    """The attention kernel derives the shifted-window mask arithmetically; the checkpoint buffers must agree."""
    for i, depth in enumerate(S.DEPTHS):
        res = S.stage_res(i)
        for b in range(depth):
            key = f"{HT}layers.{i}.blocks.{b}.attn_mask"
            if key in sd:
                if not np.array_equal(sd[key].numpy(), dsp.shifted_window_mask(res)):
                    raise RuntimeError(f"{key} differs from the standard shifted-window mask")


def check_tied_head(sd):
    if not torch.equal(sd[LMK + "lm_head.weight"], sd[LMK + "model.embed_tokens.weight"]):
        raise RuntimeError("lm_head.weight is expected to be tied to embed_tokens.weight")


def split_bf16(w):
    w = w.detach().to(torch.float32).contiguous()
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


def rope_tables(n_pos=MAX_POS):
    """cos/sin [n_pos, 32] in fp32 exactly as transformers' LlamaRotaryEmbedding computes them
    (modeling_llama.py:117-142): the two halves of the 64-wide table are identical, so 32 columns suffice."""
    inv = 1.0 / (S.ROPE_THETA ** (torch.arange(0, S.HEAD_DIM, 2, dtype=torch.int64).float() / S.HEAD_DIM))
    freqs = torch.arange(n_pos, dtype=torch.float32)[:, None] * inv[None, :]
    return freqs.cos().contiguous(), freqs.sin().contiguous()


def _rope_pair_perm():
    idx = torch.empty(S.HEAD_DIM, dtype=torch.long)
    idx[0::2] = torch.arange(0, S.HEAD_DIM // 2)
    idx[1::2] = torch.arange(S.HEAD_DIM // 2, S.HEAD_DIM)
    return idx


def build_entries(sd):
    """name -> tensor for every arena entry except the '.hi'/'.lo' split (planes are keyed by their base name)."""
    e = {}
    g = lambda k: sd[k].detach()
    real = g(HT + "spectrogram_extractor.stft.conv_real.weight")
    e["fe.window"] = real[0, 0, :].float()
    k = np.arange(512, dtype=np.float64)
    tw = np.stack([np.cos(2 * np.pi * k / S.N_FFT), -np.sin(2 * np.pi * k / S.N_FFT)], axis=1).astype(np.float32)
    e["fe.twiddle"] = torch.from_numpy(tw)
    melW = g(HT + "logmel_extractor.melW").float()
    e["fe.melW"] = melW
    nz = (melW != 0).numpy()
    lo = np.array([np.argmax(nz[:, m]) if nz[:, m].any() else 0 for m in range(S.N_MELS)], dtype=np.int32)
    hi = np.array([S.N_BINS - np.argmax(nz[::-1, m]) if nz[:, m].any() else 0 for m in range(S.N_MELS)], dtype=np.int32)
    e["fe.mel_lo"], e["fe.mel_hi"] = torch.from_numpy(lo), torch.from_numpy(hi)
    bn = HT + "bn0."
    scale = g(bn + "weight").double() / torch.sqrt(g(bn + "running_var").double() + S.BN_EPS)
    e["fe.bn_scale"] = scale.float()
    e["fe.bn_shift"] = (g(bn + "bias").double() - g(bn + "running_mean").double() * scale).float()
    e["pe.w"] = g(HT + "patch_embed.proj.weight").reshape(S.EMBED_DIM, 16)
    e["pe.b"] = g(HT + "patch_embed.proj.bias")
    e["pe.ln_w"] = g(HT + "patch_embed.norm.weight")
    e["pe.ln_b"] = g(HT + "patch_embed.norm.bias")
    for i, depth in enumerate(S.DEPTHS):
        C, nH = S.stage_dim(i), S.HEADS[i]
        qscale = (C // nH) ** -0.5
        for b in range(depth):
            src, dst = f"{HT}layers.{i}.blocks.{b}.", f"s{i}.b{b}."
            e[dst + "ln1_w"], e[dst + "ln1_b"] = g(src + "norm1.weight"), g(src + "norm1.bias")
            qkv_w, qkv_b = g(src + "attn.qkv.weight").clone(), g(src + "attn.qkv.bias").clone()
            qkv_w[:C] *= qscale
            qkv_b[:C] *= qscale
            e[dst + "qkv_w"], e[dst + "qkv_b"] = qkv_w, qkv_b
            table = g(src + "attn.relative_position_bias_table")
            index = g(src + "attn.relative_position_index").reshape(-1)
            e[dst + "relbias"] = table[index].reshape(64, 64, nH).permute(2, 0, 1).contiguous()
            e[dst + "proj_w"], e[dst + "proj_b"] = g(src + "attn.proj.weight"), g(src + "attn.proj.bias")
            e[dst + "ln2_w"], e[dst + "ln2_b"] = g(src + "norm2.weight"), g(src + "norm2.bias")
            e[dst + "fc1_w"], e[dst + "fc1_b"] = g(src + "mlp.fc1.weight"), g(src + "mlp.fc1.bias")
            e[dst + "fc2_w"], e[dst + "fc2_b"] = g(src + "mlp.fc2.weight"), g(src + "mlp.fc2.bias")
        if i < len(S.DEPTHS) - 1:
            src, dst = f"{HT}layers.{i}.downsample.", f"m{i}."
            e[dst + "ln_w"], e[dst + "ln_b"] = g(src + "norm.weight"), g(src + "norm.bias")
            e[dst + "red_w"] = g(src + "reduction.weight")
    e["tail.ln_w"], e["tail.ln_b"] = g(HT + "norm.weight"), g(HT + "norm.bias")
    # conv weight [cls, c, fb, dt] -> GEMM weight [cls, (dt*2+fb)*768 + c]
    e["tail.tscam_w"] = g(HT + "tscam_conv.weight").permute(0, 3, 2, 1).reshape(S.NUM_CLASSES, 6 * S.ENC_OUT)
    e["tail.tscam_b"] = g(HT + "tscam_conv.bias")
    c2l = torch.zeros(S.ENC_OUT, 544, dtype=torch.float32)
    c2l[:, :S.NUM_CLASSES] = g("audio_encoder.base.c2l.weight")
    e["tail.c2l_w"], e["tail.c2l_b"] = c2l, g("audio_encoder.base.c2l.bias")
    e["proj.w1"] = g("audio_encoder.projection.linear1.weight")
    e["proj.w2"] = g("audio_encoder.projection.linear2.weight")
    e["proj.ln_w"] = g("audio_encoder.projection.layer_norm.weight")
    e["proj.ln_b"] = g("audio_encoder.projection.layer_norm.bias")
    emb = g(LMK + "model.embed_tokens.weight")
    e["lm.embed"] = emb
    e["lm.head_w"] = g(LMK + "lm_head.weight")
    e["lm.norm"] = g(LMK + "model.norm.weight")
    e["lm.rope_cos"], e["lm.rope_sin"] = rope_tables()
    perm = _rope_pair_perm()
    for l in range(S.N_LAYERS):
        src, dst = f"{LMK}model.layers.{l}.", f"lm.l{l}."
        q = g(src + "self_attn.q_proj.weight").reshape(S.N_HEADS, S.HEAD_DIM, S.HIDDEN)[:, perm]
        k_ = g(src + "self_attn.k_proj.weight").reshape(S.N_KV_HEADS, S.HEAD_DIM, S.HIDDEN)[:, perm]
        v = g(src + "self_attn.v_proj.weight")
        e[dst + "qkv_w"] = torch.cat([q.reshape(-1, S.HIDDEN), k_.reshape(-1, S.HIDDEN), v], dim=0)
        e[dst + "o_w"] = g(src + "self_attn.o_proj.weight")
        gate, up = g(src + "mlp.gate_proj.weight"), g(src + "mlp.up_proj.weight")
        e[dst + "gu_w"] = torch.stack([gate, up], dim=1).reshape(2 * S.INTER, S.HIDDEN)
        e[dst + "down_w"] = g(src + "mlp.down_proj.weight")
        e[dst + "ln1"] = g(src + "input_layernorm.weight")
        e[dst + "ln2"] = g(src + "post_attention_layernorm.weight")
    return e


def pack(sd, entries, total_bytes, verify=True):
    """Return a CPU uint8 tensor of ``total_bytes`` holding the arena for ``entries`` [(name, offset, bytes)]."""
    sd = strip_module_prefix(sd)
    if verify:
        check_schema(sd)
        check_frontend_basis(sd)
        check_shift_masks(sd)
        check_tied_head(sd)
    src = build_entries(sd)
    arena = torch.zeros(total_bytes, dtype=torch.uint8)
    planes = {}
    used = set()
    for name, offset, nbytes in entries:
        if name.endswith(".hi") or name.endswith(".lo"):
            base = name[:-3]
            if base not in planes:
                planes[base] = split_bf16(src[base])
            t = planes[base][0 if name.endswith(".hi") else 1]
            used.add(base)
        else:
            t = src[name]
            t = t.to(torch.int32) if t.dtype in (torch.int32, torch.int64) else t.to(torch.float32)
            used.add(name)
        raw = t.contiguous().view(-1).view(torch.uint8)
        if raw.numel() != nbytes:
            raise RuntimeError(f"arena entry {name}: packer produced {raw.numel()} bytes, library expects {nbytes}")
        arena[offset:offset + nbytes] = raw
    unused = set(src) - used
    if unused:
        raise RuntimeError(f"packer entries not consumed by the library layout: {sorted(unused)[:5]}")
    return arena
