"""Thin Python face of the C-ABI: one ``Engine`` = one ``mb_create`` handle on one GPU.

PyTorch tensors are only containers for device memory here (allocation, H2D copies, the NCCL broadcast of the
weight arena); every computation is a call into ``libmellow_b200.so``.  There is no fallback path: constructing an
``Engine`` without the library or without a CUDA device raises.
"""
import ctypes

import torch

from . import native, schema as S, weights

POLICY_SPLIT, POLICY_FAST, POLICY_SPLIT24 = 0, 1, 2
POLICIES = {"split": POLICY_SPLIT, "bf16x3": POLICY_SPLIT, "fast": POLICY_FAST, "bf16": POLICY_FAST,
            "split24": POLICY_SPLIT24}


class MellowNativeError(RuntimeError):
    pass


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _LogitsOut:
    """What ``lm(inputs_embeds=...)`` returns: ``.logits`` of shape (B, 1, V) -- the last position only."""
    def __init__(self, logits):
        self.logits = logits


class _LmShim:
    """Attribute-compatible stand-in for ``model.caption_decoder.lm`` (a transformers ``LlamaForCausalLM`` in the
    reference): callable with ``inputs_embeds=`` and exposing ``.model.embed_tokens`` (wrapper.py:217,237)."""
    def __init__(self, engine):
        self._engine = engine
        self.model = self
        self.lm = self                                   # so that `engine.caption_decoder.lm` resolves to this object

    def __call__(self, inputs_embeds=None, **unused):
        if inputs_embeds is None:
            raise ValueError("only lm(inputs_embeds=...) is supported, like the reference call at wrapper.py:217")
        return _LogitsOut(self._engine.lm_forward_last(inputs_embeds)[:, None, :])

    def embed_tokens(self, ids):
        return self._engine.embed_tokens(ids)


class Engine:
    def __init__(self, state_dict=None, device=0, max_batch=8, max_new_tokens=300, policy="split", arena=None,
                 verify=True):
        if not torch.cuda.is_available():
            raise MellowNativeError("mellow_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = native.load()
        self.device = torch.device("cuda", device)
        self.max_batch, self.max_new_tokens = int(max_batch), int(max_new_tokens)
        self.policy = POLICIES[policy] if isinstance(policy, str) else int(policy)
        self.handle = self.lib.mb_create(device, self.max_batch, self.max_new_tokens, self.policy)
        if not self.handle:
            raise MellowNativeError("mb_create failed: " + self.lib.mb_last_error(None).decode())
        self.arena = None
        self.caption_decoder = _LmShim(self)             # reference seam: model.caption_decoder.lm(...)
        if arena is not None:
            self.bind_arena(arena)
        elif state_dict is not None:
            self.bind_arena(self.pack_arena(state_dict, verify=verify).to(self.device))

    # ------------------------------------------------------------------ weights
    def pack_arena(self, state_dict, verify=True):
        return weights.pack(state_dict, native.weight_entries(self.lib), self.lib.mb_weights_size(), verify=verify)

    def arena_bytes(self):
        return self.lib.mb_weights_size()

    def bind_arena(self, arena):
        assert arena.is_cuda and arena.dtype == torch.uint8 and arena.numel() == self.lib.mb_weights_size()
        self.arena = arena
        self._ck(self.lib.mb_bind_weights(self.handle, _ptr(arena), arena.numel()))

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        if rc != 0:
            raise MellowNativeError(self.lib.mb_last_error(self.handle).decode())

    def _stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        return ctypes.c_void_p(s)     # 0 (legacy default stream) makes the library use its own blocking stream

    def _dev(self, t, dtype):
        return t.to(device=self.device, dtype=dtype).contiguous()

    def eval(self):
        """The reference calls ``model.eval()`` (wrapper.py:49,205); inference is the only mode here."""
        return self

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_launches(self):
        return self.lib.mb_kernel_launches(self.handle)

    def set_option(self, name, value):
        """Handle-scoped options of ``mb_set_option`` (include/mellow_b200.h): 'graph', 'decode_unfused',
        'skip_finished', 'kv_prefetch', 'gemm_engine' (lab builds)."""
        self._ck(self.lib.mb_set_option(self.handle, name.encode(), int(value)))

    def workspace_bytes(self):
        return self.lib.mb_workspace_bytes(self.handle)

    # ------------------------------------------------------------------ stages
    def frontend(self, wave, want_logmel=True, want_bn=True):
        wave = self._dev(wave, torch.float32)
        n = wave.shape[0]
        assert wave.shape[1] == S.CLIP_SAMPLES
        lm = torch.empty(n, S.N_FRAMES, S.N_MELS, device=self.device) if want_logmel else None
        bn = torch.empty(n, S.N_FRAMES, S.N_MELS, device=self.device) if want_bn else None
        self._ck(self.lib.mb_frontend(self.handle, _ptr(wave), n, _ptr(lm), _ptr(bn), self._stream()))
        return lm, bn

    def encode(self, wave1, wave2):
        wave1, wave2 = self._dev(wave1, torch.float32), self._dev(wave2, torch.float32)
        b = wave1.shape[0]
        rows = torch.empty(2, b, S.AUDIO_FRAMES + 1, S.D_PROJ, device=self.device)
        self._ck(self.lib.mb_encode(self.handle, _ptr(wave1), _ptr(wave2), b, _ptr(rows), self._stream()))
        return rows

    def encode_heads(self, batch, expand=True):
        """the output dicts the reference returns as od1 / od2 (mellow.py:100-108) for the last
        encode()/generate() of `batch` pairs.  expand=True gives the reference shapes (framewise (B,1024,527),
        embedding (B,1025,768)) as views that repeat each of the 32 unique frame rows 32 times (htsat.py:43-56)."""
        n = 2 * batch
        clip = torch.empty(n, S.NUM_CLASSES, device=self.device)
        frame = torch.empty(n, S.AUDIO_FRAMES, S.NUM_CLASSES, device=self.device)
        latent = torch.empty(n, S.ENC_OUT, device=self.device)
        femb = torch.empty(n, S.AUDIO_FRAMES, S.ENC_OUT, device=self.device)
        self._ck(self.lib.mb_encode_heads(self.handle, n, _ptr(clip), _ptr(frame), _ptr(latent), _ptr(femb), self._stream()))
        out = []
        for k in range(2):
            sl = slice(k * batch, (k + 1) * batch)
            fw, fe = frame[sl], femb[sl]
            if expand:
                fw = fw[:, :, None, :].expand(-1, -1, 32, -1).reshape(batch, 32 * S.AUDIO_FRAMES, S.NUM_CLASSES)
                fe = fe[:, :, None, :].expand(-1, -1, 32, -1).reshape(batch, 32 * S.AUDIO_FRAMES, S.ENC_OUT)
                emb = torch.cat([latent[sl][:, None, :], fe], dim=1)
            else:
                emb = torch.cat([latent[sl][:, None, :], fe], dim=1)
            out.append({"framewise_output": fw, "clipwise_output": clip[sl], "latent_output": latent[sl], "embedding": emb})
        return out[0], out[1]

    def generate_prefix_inference(self, input_dict, heads=True):
        """The reference's inner seam ``model.generate_prefix_inference`` (mellow.py:100-108):
        {'audio1': (B,320000), 'audio2': (B,320000), 'input': {'input_ids': (B,129)}} -> (prefix (B,389,576), od1, od2)."""
        a1, a2 = input_dict["audio1"], input_dict["audio2"]
        self.encode(a1, a2)
        prefix = self.prefix(input_dict["input"]["input_ids"])
        od1, od2 = self.encode_heads(a1.shape[0]) if heads else (None, None)
        return prefix, od1, od2

    _TAP_SHAPES = {0: (4096, 96), 1: (1024, 192), 2: (256, 384), 3: (64, 768), 4: (64, 768), 5: (33, 768)}

    def encode_tap(self, wave, stage):
        wave = self._dev(wave, torch.float32)
        n = wave.shape[0]
        r, c = self._TAP_SHAPES[stage]
        out = torch.empty(n * r * c, device=self.device)
        self._ck(self.lib.mb_encode_tap(self.handle, _ptr(wave), n, stage, _ptr(out), self._stream()))
        if stage == 5:
            return out[:n * 768].view(n, 768), out[n * 768:].view(n, 32, 768)
        return out.view(n, r, c)

    def prefix(self, input_ids):
        ids = self._dev(input_ids, torch.int32)
        b = ids.shape[0]
        out = torch.empty(b, S.PREFIX_LEN, S.HIDDEN, device=self.device)
        self._ck(self.lib.mb_prefix(self.handle, _ptr(ids), b, _ptr(out), self._stream()))
        return out

    def set_prefix(self, prefix):
        prefix = self._dev(prefix, torch.float32)
        assert tuple(prefix.shape[1:]) == (S.PREFIX_LEN, S.HIDDEN)
        self._ck(self.lib.mb_set_prefix(self.handle, _ptr(prefix), prefix.shape[0], self._stream()))
        torch.cuda.synchronize(self.device)

    def prefill(self, batch, want_logits=True):
        out = torch.empty(batch, S.VOCAB, device=self.device) if want_logits else None
        self._ck(self.lib.mb_prefill(self.handle, batch, _ptr(out), self._stream()))
        return out

    def lm_forward_last(self, inputs_embeds):
        """``model.caption_decoder.lm(inputs_embeds=x).logits[:, -1]`` of the reference (wrapper.py:217-218): one
        cache-less causal forward over (B,S,576) embeddings -> last-position logits (B,49152)."""
        x = self._dev(inputs_embeds, torch.float32)
        b, s_len, hid = x.shape
        assert hid == S.HIDDEN
        out = torch.empty(b, S.VOCAB, device=self.device)
        self._ck(self.lib.mb_lm_forward_last(self.handle, _ptr(x), b, s_len, _ptr(out), self._stream()))
        return out

    def embed_tokens(self, ids):
        """``lm.model.embed_tokens(ids)`` (wrapper.py:237): integer ids of any shape -> (..., 576) float32."""
        flat = self._dev(ids, torch.int32).reshape(-1)
        out = torch.empty(flat.numel(), S.HIDDEN, device=self.device)
        self._ck(self.lib.mb_embed_tokens(self.handle, _ptr(flat), flat.numel(), _ptr(out), self._stream()))
        return out.view(*ids.shape, S.HIDDEN)

    def decode(self, batch, max_len, temperature=1.0, top_p=0.8, eos_id=0, dump_logits=False, forced_tokens=None):
        toks = torch.empty(batch, max_len, dtype=torch.int32, device=self.device)
        dump = torch.empty(max_len, batch, S.VOCAB, device=self.device) if dump_logits else None
        forced = self._dev(forced_tokens, torch.int32) if forced_tokens is not None else None
        steps = ctypes.c_int(0)
        self._ck(self.lib.mb_decode(self.handle, batch, max_len, temperature, top_p, eos_id, _ptr(toks),
                                    ctypes.byref(steps), _ptr(dump), _ptr(forced), self._stream()))
        toks = toks[:, :steps.value]
        return (toks, dump[:steps.value]) if dump_logits else toks

    def generate(self, wave1, wave2, input_ids, max_len, temperature=1.0, top_p=0.8, eos_id=0):
        """Device-resident inputs -> tokens (B, steps) int32 on device."""
        wave1, wave2 = self._dev(wave1, torch.float32), self._dev(wave2, torch.float32)
        ids = self._dev(input_ids, torch.int32)
        b = wave1.shape[0]
        toks = torch.empty(b, max_len, dtype=torch.int32, device=self.device)
        steps = ctypes.c_int(0)
        self._ck(self.lib.mb_generate(self.handle, _ptr(wave1), _ptr(wave2), _ptr(ids), b, max_len, temperature,
                                      top_p, eos_id, _ptr(toks), ctypes.byref(steps), self._stream()))
        return toks[:, :steps.value]

    def generate_host(self, wave1, wave2, input_ids, max_len, temperature=1.0, top_p=0.8, eos_id=0, out=None):
        """HOST tensors in (pinned recommended), HOST int32 tokens out; the copies happen inside the call."""
        assert not wave1.is_cuda and not wave2.is_cuda and not input_ids.is_cuda
        wave1, wave2 = wave1.contiguous(), wave2.contiguous()
        ids = input_ids.to(torch.int32).contiguous()
        b = wave1.shape[0]
        if out is None:
            out = torch.empty(b, max_len, dtype=torch.int32).pin_memory()
        steps = ctypes.c_int(0)
        self._ck(self.lib.mb_generate_host(self.handle, _ptr(wave1), _ptr(wave2), _ptr(ids), b, max_len, temperature,
                                           top_p, eos_id, _ptr(out), ctypes.byref(steps), self._stream()))
        return out[:, :steps.value]

    # ------------------------------------------------------------------ audio ingest
    def _resample_kernel(self, sr_in, sr_out):
        key = (int(sr_in), int(sr_out))
        cache = self.__dict__.setdefault("_rs_cache", {})
        if key not in cache:
            from .audio_io import sinc_resample_kernel
            kern, width, orig, new = sinc_resample_kernel(sr_in, sr_out)
            cache[key] = (kern.contiguous().to(self.device), width, orig, new)
        return cache[key]

    def prepare_clip(self, pcm, sample_rate, target_rate=S.SAMPLE_RATE, resample=True, rng=None, out=None):
        """pcm (channels, n) float32 host/device tensor -> (320000,) device tensor: per-channel polyphase resampling to
        `target_rate`, channels concatenated (reference wrapper.py:146-149), then tile-or-crop (wrapper.py:152-167)."""
        import random as _random
        from .audio_io import plan_fit, resampled_length
        rng = rng or _random
        pcm = self._dev(pcm if pcm.dim() == 2 else pcm[None, :], torch.float32)
        ch, n_in = pcm.shape
        if resample and int(sample_rate) != int(target_rate):
            kern, width, orig, new = self._resample_kernel(sample_rate, target_rate)
            n_out = resampled_length(n_in, orig, new)
            flat = torch.empty(ch * n_out, device=self.device)
            for c in range(ch):
                self._ck(self.lib.mb_audio_resample(self.handle, _ptr(pcm[c]), n_in, orig, new, _ptr(kern), kern.shape[1],
                                                    width, ctypes.c_void_p(flat.data_ptr() + 4 * c * n_out), n_out,
                                                    self._stream()))
        else:
            flat = pcm.reshape(-1)
        total = flat.numel()
        start = plan_fit(total, S.CLIP_SAMPLES, rng)
        if out is None:
            out = torch.empty(S.CLIP_SAMPLES, device=self.device)
        self._ck(self.lib.mb_audio_fit(self.handle, _ptr(flat), total, start, _ptr(out), self._stream()))
        return out

    def bench_decode_attention(self, batch, ctx, iters):
        self._ck(self.lib.mb_bench_decode_attention(self.handle, batch, ctx, iters, self._stream()))

    def op_gemm(self, a, w, bias=None, act=0):
        a, w = self._dev(a, torch.float32), self._dev(w, torch.float32)
        bias = self._dev(bias, torch.float32) if bias is not None else None
        m, k = a.shape
        n = w.shape[0]
        c = torch.empty(m, n, device=self.device)
        self._ck(self.lib.mb_op_gemm(self.handle, _ptr(a), _ptr(w), _ptr(bias), _ptr(c), m, n, k, act, self._stream()))
        return c
