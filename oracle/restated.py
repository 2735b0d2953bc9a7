"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path
(`mellow_b200/`); only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline`
/ `--impl reference` legs of `bench.py` may use it, and only as the checker.

Standalone CPU restatement (plain PyTorch fp32 ops, no reference imports) of the
reference's two-audio-plus-prompt inference path, operating directly on a flat
checkpoint ``state_dict`` with the reference's schema.  It exists because
``/root/reference`` does not travel to the GPU box; it is *pinned* against the
reference's own model classes by ``tests/test_oracle_pinned.py`` (runs wherever
/root/reference is present) and against the committed golden vectors produced by
those classes (``tests/golden/make_golden.py``).

PARITY PIN STATUS: the reference repo has no tests, golden vectors or expected
outputs -- "parity unpinned" by the reference's own
fixtures.  The pins used instead are outputs of the reference's own model code
executed in the build container and committed under ``tests/golden/``.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Third-party arithmetic that the reference delegates to
packages outside its tree is restated from the published algorithms:
``torchlibrosa==0.1.0`` (Spectrogram / LogmelFilterBank) and
``transformers`` ``LlamaForCausalLM`` (SmolLM2-135M), see SURVEY.md Appendix A.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------- constants
# mellow/model/config.py:4-9, mellow/config/v0.yaml, mellow/model/htsat.py:599-606
N_FFT, HOP, N_MELS = 1024, 320, 64
SPEC_SIZE, FREQ_RATIO, PATCH, WINDOW = 256, 4, 4, 8
DEPTHS, HEADS, EMBED = (2, 2, 6, 2), (4, 8, 16, 32), 96
VOCAB, HIDDEN, N_LAYERS, N_HEADS, N_KV, HEAD_DIM = 49152, 576, 30, 9, 3, 64
RMS_EPS, ROPE_THETA = 1e-5, 100000.0
HT = "audio_encoder.base.htsat."
LM = "caption_decoder.lm."


# ----------------------------------------------------------------------------- front end
def spectrogram(sd, wave):
    """torchlibrosa Spectrogram(center=True, reflect, power=2) as called at mellow/model/htsat.py:864
    (constructed :647-649): reflect-pad 512, two conv1d with the checkpoint's DFT bases, re^2+im^2.
    wave (N,320000) -> (N,1001,513)."""
    x = F.pad(wave[:, None, :], (N_FFT // 2, N_FFT // 2), mode="reflect")
    re = F.conv1d(x, sd[HT + "spectrogram_extractor.stft.conv_real.weight"], stride=HOP)
    im = F.conv1d(x, sd[HT + "spectrogram_extractor.stft.conv_imag.weight"], stride=HOP)
    return (re * re + im * im).transpose(1, 2)


def logmel(sd, spec):
    """torchlibrosa LogmelFilterBank(ref=1, amin=1e-10, top_db=None) as called at htsat.py:865
    (constructed :651-653). (N,1001,513) -> (N,1001,64)."""
    mel = torch.matmul(spec, sd[HT + "logmel_extractor.melW"])
    return 10.0 * torch.log10(torch.clamp(mel, min=1e-10))


def bn0(sd, x):
    """Eval-mode BatchNorm2d over the mel axis, htsat.py:868-870. (N,1001,64) -> same."""
    p = HT + "bn0."
    inv = torch.rsqrt(sd[p + "running_var"] + 1e-5)
    return (x - sd[p + "running_mean"]) * inv * sd[p + "weight"] + sd[p + "bias"]


def cubic_weights(t, A=-0.75):
    """Cubic-convolution coefficients (ATen upsample_bicubic2d, A=-0.75) for fractional offset t."""
    def c1(x):
        return ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0
    def c2(x):
        return ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A
    return torch.stack([c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)], dim=-1)


def stretch_time(x, target=SPEC_SIZE * FREQ_RATIO):
    """F.interpolate(mode='bicubic', align_corners=True) from T=1001 to 1024 along time, identity along
    mel (htsat.py:836-837); restated as the 1-D cubic convolution it equals.
    (N,T,64) -> (N,1024,64)."""
    n, t_in, f = x.shape
    scale = (t_in - 1) / (target - 1)
    src = torch.arange(target, dtype=torch.float32) * scale
    i0 = torch.floor(src)
    w = cubic_weights(src - i0)                                   # (1024,4)
    idx = (i0.long()[:, None] + torch.arange(-1, 3)[None, :]).clamp(0, t_in - 1)   # (1024,4)
    g = x[:, idx, :]                                              # (N,1024,4,64)
    return (g * w[None, :, :, None]).sum(dim=2)


def fold_image(x):
    """Second half of reshape_wav2img, htsat.py:840-844: img[c*64+f, t] = x[c*256+t, f].
    (N,1024,64) -> (N,256,256)."""
    n = x.shape[0]
    x = x.transpose(1, 2)                                         # (N,64,1024)
    x = x.reshape(n, N_MELS, FREQ_RATIO, SPEC_SIZE)               # (N,f,c,t)
    return x.permute(0, 2, 1, 3).reshape(n, SPEC_SIZE, SPEC_SIZE)


# ----------------------------------------------------------------------------- Swin encoder
def patch_embed(sd, img):
    """PatchEmbed.forward htsat.py:108-116: conv 4x4 stride 4 (1->96) + bias, flatten, LayerNorm(96)."""
    x = F.conv2d(img[:, None], sd[HT + "patch_embed.proj.weight"], sd[HT + "patch_embed.proj.bias"], stride=PATCH)
    x = x.flatten(2).transpose(1, 2)
    return F.layer_norm(x, (EMBED,), sd[HT + "patch_embed.norm.weight"], sd[HT + "patch_embed.norm.bias"], 1e-5)


def _partition(x, ws):
    b, h, w, c = x.shape                                          # htsat.py:224-235
    x = x.view(b, h // ws, ws, w // ws, ws, c)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, c)


def _reverse(win, ws, h, w):
    b = win.shape[0] // ((h // ws) * (w // ws))                   # htsat.py:238-251
    x = win.view(b, h // ws, w // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, -1)


def window_attention(sd, p, xw, n_heads, mask):
    """WindowAttention.forward htsat.py:301-332 (the returned attention map is not needed)."""
    b_, n, c = xw.shape
    hd = c // n_heads
    qkv = F.linear(xw, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    qkv = qkv.reshape(b_, n, 3, n_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    table = sd[p + "attn.relative_position_bias_table"]
    index = sd[p + "attn.relative_position_index"].view(-1)
    attn = attn + table[index].view(n, n, n_heads).permute(2, 0, 1).unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        attn = attn.view(b_ // nw, nw, n_heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, n_heads, n, n)
    attn = torch.softmax(attn, dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(b_, n, c)
    return F.linear(out, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])


def swin_block(sd, p, x, res, n_heads, shift):
    """SwinTransformerBlock.forward htsat.py:414-455 (eval: drop_path identity)."""
    b, l, c = x.shape
    shortcut = x
    y = F.layer_norm(x, (c,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5).view(b, res, res, c)
    if shift > 0:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
    yw = _partition(y, WINDOW)
    mask = sd[p + "attn_mask"] if shift > 0 else None
    aw = window_attention(sd, p, yw, n_heads, mask)
    y = _reverse(aw, WINDOW, res, res)
    if shift > 0:
        y = torch.roll(y, shifts=(shift, shift), dims=(1, 2))
    x = shortcut + y.reshape(b, l, c)
    z = F.layer_norm(x, (c,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    z = F.linear(z, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    z = F.gelu(z)                                                 # nn.GELU() exact erf, htsat.py:121,132
    z = F.linear(z, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + z


def patch_merging(sd, p, x, res):
    """PatchMerging.forward htsat.py:478-499."""
    b, l, c = x.shape
    x = x.view(b, res, res, c)
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], dim=-1)
    x = x.view(b, -1, 4 * c)
    x = F.layer_norm(x, (4 * c,), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    return F.linear(x, sd[p + "reduction.weight"])


def swin_stages(sd, x, taps=None):
    """BasicLayer loop htsat.py:739-740, 553-567 (the dead eval-mode attention mean :560-566 is skipped)."""
    for i, depth in enumerate(DEPTHS):
        res = (SPEC_SIZE // PATCH) >> i
        for b in range(depth):
            shift = 0 if (b % 2 == 0 or res <= WINDOW) else WINDOW // 2          # htsat.py:539, 368-371
            x = swin_block(sd, f"{HT}layers.{i}.blocks.{b}.", x, res, HEADS[i], shift)
        if i < len(DEPTHS) - 1:
            x = patch_merging(sd, f"{HT}layers.{i}.downsample.", x, res)
        if taps is not None:
            taps[f"stage{i}"] = x
    return x


def encoder_tail(sd, x):
    """forward_features TSCAM branch htsat.py:742-796 + HTSATWrapper.forward :950-955, keeping only the
    33 unique rows per clip: returns latent (N,768) and the 32 unique c2l(framewise) rows (N,32,768).
    The reference repeats every frame row 32x (interpolate :43-56, :780) before c2l."""
    n = x.shape[0]
    x = F.layer_norm(x, (768,), sd[HT + "norm.weight"], sd[HT + "norm.bias"], 1e-5)       # :744
    x = x.permute(0, 2, 1).reshape(n, 768, 8, 8)                                          # :748 (B,C,SF,ST)
    x = x.reshape(n, 768, 4, 2, 8).permute(0, 1, 3, 2, 4).reshape(n, 768, 2, 32)          # :752-753
    latent = torch.flatten(x, 2).mean(dim=-1)                                             # :756-757
    y = F.conv2d(x, sd[HT + "tscam_conv.weight"], sd[HT + "tscam_conv.bias"], padding=(0, 1))   # :774
    frames = torch.sigmoid(torch.flatten(y, 2)).permute(0, 2, 1)                          # :775,780 (N,32,527)
    oframe = F.linear(frames, sd["audio_encoder.base.c2l.weight"], sd["audio_encoder.base.c2l.bias"])   # :952
    return latent, oframe


def encoder_heads(sd, x):
    """The classification outputs of the same TSCAM branch (htsat.py:774-796, loss_type "clip_bce" config.py:1), which
    the reference returns inside od1/od2 (mellow.py:100-108) and generate() ignores: clipwise (N,527) =
    sigmoid(avgpool_t(conv)), and the 32 unique framewise rows (N,32,527) = sigmoid(conv) (repeated 32x by :780)."""
    n = x.shape[0]
    x = F.layer_norm(x, (768,), sd[HT + "norm.weight"], sd[HT + "norm.bias"], 1e-5)
    x = x.permute(0, 2, 1).reshape(n, 768, 8, 8)
    x = x.reshape(n, 768, 4, 2, 8).permute(0, 1, 3, 2, 4).reshape(n, 768, 2, 32)
    y = torch.flatten(F.conv2d(x, sd[HT + "tscam_conv.weight"], sd[HT + "tscam_conv.bias"], padding=(0, 1)), 2)   # (N,527,32)
    framewise = torch.sigmoid(y).permute(0, 2, 1)                                         # :780 before interpolate
    clipwise = torch.sigmoid(y.mean(dim=-1))                                              # :782-783,795
    return clipwise, framewise


def projection(sd, x):
    """Projection.forward mellow/model/mellow.py:48-52 (dropout is eval-identity)."""
    e1 = F.linear(x, sd["audio_encoder.projection.linear1.weight"])
    e2 = F.linear(F.gelu(e1), sd["audio_encoder.projection.linear2.weight"])
    return F.layer_norm(e1 + e2, (576,), sd["audio_encoder.projection.layer_norm.weight"],
                        sd["audio_encoder.projection.layer_norm.bias"], 1e-5)


def encode_clips(sd, wave, taps=None):
    """AudioEncoder.forward mellow.py:64-68 restricted to the unique rows: (N,320000) -> (N,33,576)
    = [projected latent, 32 projected frame rows]."""
    spec = spectrogram(sd, wave)
    lm_ = logmel(sd, spec)
    xb = bn0(sd, lm_)
    img = fold_image(stretch_time(xb))
    tok = patch_embed(sd, img)
    if taps is not None:
        taps.update(logmel=lm_, bn=xb, image=img, patch=tok)
    x = swin_stages(sd, tok, taps)
    latent, oframe = encoder_tail(sd, x)
    emb33 = torch.cat([latent[:, None, :], oframe], dim=1)                                # (N,33,768)
    out = projection(sd, emb33)
    if taps is not None:
        taps.update(latent=latent, oframe=oframe, rows33=out, final_tokens=x)
    return out


def expand_audio_rows(rows33):
    """downsample() mellow/model/decoder.py:14-18 applied to the reference's (N,1025,576) tensor whose
    rows 1.. are the 32 frame rows each repeated 32x: avg-pool by 8 gives each frame row 4 times.
    The mean of 8 identical floats is reproduced with the same fp32 summation (sum of 8, times 1/8)."""
    lat, fr = rows33[:, :1], rows33[:, 1:]
    rep = fr[:, :, None, :].expand(-1, -1, 32, -1).reshape(fr.shape[0], 1024, fr.shape[-1])
    pooled = F.avg_pool2d(rep, kernel_size=(8, 1))
    return torch.cat([lat, pooled], dim=1)                                                # (N,129,576)


def build_prefix(sd, rows33_a, rows33_b, input_ids):
    """DecoderModel.generate_prefix_inference decoder.py:36-55 (smollm2 branch): sep token id 0."""
    emb = sd[LM + "model.embed_tokens.weight"]
    b = input_ids.shape[0]
    sep = emb[0][None, None, :].expand(b, 1, -1)
    return torch.cat([expand_audio_rows(rows33_a), sep, expand_audio_rows(rows33_b), sep, emb[input_ids]], dim=1)


# ----------------------------------------------------------------------------- SmolLM2 (Llama) forward
def rms_norm(x, w, eps=RMS_EPS):
    """transformers LlamaRMSNorm (modeling_llama.py:62-67), fp32."""
    return w * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))


def rope_tables(seq_len):
    """LlamaRotaryEmbedding default rope (modeling_llama.py:117-142): cos/sin (S,64)."""
    inv = 1.0 / (ROPE_THETA ** (torch.arange(0, HEAD_DIM, 2, dtype=torch.int64).float() / HEAD_DIM))
    freqs = torch.arange(seq_len, dtype=torch.float32)[:, None] * inv[None, :]
    emb = torch.cat([freqs, freqs], dim=-1)
    return emb.cos(), emb.sin()


def _rot_half(x):
    return torch.cat([-x[..., HEAD_DIM // 2:], x[..., :HEAD_DIM // 2]], dim=-1)


def round_to_24_bits(t):
    """What mellow_b200's policy split24 stores in its KV cache: fp32 rounded (nearest-even) to sign + 8 exponent + 15
    mantissa bits.  Test infrastructure for the tolerance claim of that policy; the reference keeps fp32."""
    u = t.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    r = (u + 0x7F + ((u >> 8) & 1)) & 0xFFFFFF00
    r = torch.where(r >= 2 ** 31, r - 2 ** 32, r).to(torch.int32)
    return r.view(torch.float32).view_as(t)


def llama_hidden(sd, x, taps=None, kv_round=None):
    """LlamaModel.forward with inputs_embeds, no cache, no attention mask => pure causal mask over all
    positions, pads attended (reference wrapper.py:217; transformers modeling_llama.py:375-425).
    x (B,S,576) -> final-normed hidden (B,S,576).  kv_round (optional, not in the reference): applied to the roped keys
    and to the values, i.e. to what a KV cache would hold."""
    b, s, _ = x.shape
    cos, sin = rope_tables(s)
    causal = torch.full((s, s), float("-inf")).triu(1)
    for l in range(N_LAYERS):
        p = f"{LM}model.layers.{l}."
        h = rms_norm(x, sd[p + "input_layernorm.weight"])
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"]).view(b, s, N_HEADS, HEAD_DIM).transpose(1, 2)
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"]).view(b, s, N_KV, HEAD_DIM).transpose(1, 2)
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"]).view(b, s, N_KV, HEAD_DIM).transpose(1, 2)
        q = q * cos + _rot_half(q) * sin                                               # :166-167
        k = k * cos + _rot_half(k) * sin
        if kv_round is not None:
            k, v = kv_round(k), kv_round(v)
        k = k.repeat_interleave(N_HEADS // N_KV, dim=1)                                # repeat_kv :187-196
        v = v.repeat_interleave(N_HEADS // N_KV, dim=1)
        att = (q @ k.transpose(-2, -1)) * (HEAD_DIM ** -0.5) + causal
        att = torch.softmax(att, dim=-1, dtype=torch.float32)
        o = (att @ v).transpose(1, 2).reshape(b, s, N_HEADS * HEAD_DIM)
        x = x + F.linear(o, sd[p + "self_attn.o_proj.weight"])
        h = rms_norm(x, sd[p + "post_attention_layernorm.weight"])
        g = F.linear(h, sd[p + "mlp.gate_proj.weight"])
        u = F.linear(h, sd[p + "mlp.up_proj.weight"])
        x = x + F.linear(F.silu(g) * u, sd[p + "mlp.down_proj.weight"])               # :182-184
        if taps is not None and l in (0, N_LAYERS - 1):
            taps[f"lm_layer{l}"] = x
    return rms_norm(x, sd[LM + "model.norm.weight"])


def last_logits(sd, hidden):
    """lm_head on the last position only (the reference computes all positions and slices, wrapper.py:217-219)."""
    return F.linear(hidden[:, -1, :], sd[LM + "lm_head.weight"])


def top_p_argmax(logits, top_p, temperature):
    """The reference's 'sampling' step wrapper.py:219-232: temperature, sort, softmax, cumsum, shifted
    top-p mask, -inf scatter, then ARGMAX (deterministic).  logits (B,V) -> (next (B,), scaled logits)."""
    logits = logits / (temperature if temperature > 0 else 1.0)
    kept = logits.clone()
    s_logits, s_idx = torch.sort(kept, descending=True)
    cum = torch.cumsum(F.softmax(s_logits, dim=-1), dim=-1)
    remove = cum > top_p
    remove[..., 1:] = remove[..., :-1].clone()
    remove[..., 0] = 0
    for r in range(kept.shape[0]):
        kept[r, s_idx[r][remove[r]]] = -float("inf")
    return torch.argmax(kept, -1), logits


@torch.no_grad()
def generate_ids(sd, prefix, max_len, top_p=0.8, temperature=1.0, stop_id=0, dump_logits=False):
    """_generate_batch wrapper.py:197-256 without tqdm / detokenisation: cache-less loop, stop when every
    row has emitted ``stop_id`` at least once.  Returns tokens (B,steps) int64 (and per-step last logits)."""
    emb = sd[LM + "model.embed_tokens.weight"]
    generated, tokens, dumped = prefix, None, []
    for _ in range(max_len):
        logits = last_logits(sd, llama_hidden(sd, generated))
        nxt, scaled = top_p_argmax(logits, top_p, temperature)
        if dump_logits:
            dumped.append(scaled)
        nxt = nxt[:, None]
        tokens = nxt if tokens is None else torch.cat([tokens, nxt], dim=1)
        generated = torch.cat([generated, emb[nxt]], dim=1)
        if ((tokens == stop_id).sum(dim=-1) > 0).all():
            break
    return (tokens, dumped) if dump_logits else tokens


@torch.no_grad()
def generate_from_wave(sd, wave1, wave2, input_ids, max_len, top_p=0.8, temperature=1.0, dump_logits=False):
    """Mellow.generate_prefix_inference mellow.py:100-108 + _generate_batch, from prepared tensors."""
    prefix = build_prefix(sd, encode_clips(sd, wave1), encode_clips(sd, wave2), input_ids)
    return generate_ids(sd, prefix, max_len, top_p, temperature, dump_logits=dump_logits)


# ----------------------------------------------------------------------------- host-side audio / text prep
def tile_or_crop(samples, target, rng):
    """load_audio_into_tensor wrapper.py:152-168 after resampling: tile+truncate when short (>= branch), random
    crop with ``rng.randrange`` when long.  ``samples`` 1-D float tensor."""
    n = samples.shape[0]
    if target >= n:
        rep = int(math.ceil(target / n))
        return samples.repeat(rep)[:target]
    start = rng.randrange(n - target)
    return samples[start:start + target]


def trim_at_stop(ids, stop_id=0):
    """Detokenisation cut wrapper.py:254: text before the first stop token."""
    out = []
    for t in ids:
        if int(t) == stop_id:
            break
        out.append(int(t))
    return out
