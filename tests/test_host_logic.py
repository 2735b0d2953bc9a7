"""CPU: host-side logic of the product package and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from mellow_b200 import dsp, native, schema as S, synth, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return native.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "mellow_b200.h")).read()
    declared = set(re.findall(r"\b(mb_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    bound = {name for name, _, _ in native.SYMBOLS}
    assert declared == bound, f"header/ctypes mismatch: {declared ^ bound}"
    for name in declared:
        assert hasattr(lib, name)
    assert lib.mb_version() >= 1


def test_arena_layout_is_consistent(lib):
    entries = native.weight_entries(lib)
    assert len({n for n, _, _ in entries}) == len(entries)
    end = 0
    for name, off, nbytes in entries:
        assert off % 256 == 0 and off >= end and nbytes > 0, name
        end = off + nbytes
    assert end <= lib.mb_weights_size()


def test_pack_roundtrip(sd, lib):
    entries = native.weight_entries(lib)
    arena = weights.pack(sd, entries, lib.mb_weights_size())
    table = {n: (o, b) for n, o, b in entries}

    def f32(name):
        o, b = table[name]
        return arena[o:o + b].view(torch.float32)

    def planes(name):
        o, b = table[name + ".hi"]
        hi = arena[o:o + b].view(torch.bfloat16).float()
        o, b = table[name + ".lo"]
        return hi + arena[o:o + b].view(torch.bfloat16).float()

    # hi + lo reproduces fp32 weights to ~2^-16 relative
    w = sd["caption_decoder.lm.model.layers.3.self_attn.o_proj.weight"]
    assert (planes("lm.l3.o_w").view_as(w) - w).abs().max() <= w.abs().max() * 2.0 ** -15
    # q rows of the Swin qkv are pre-scaled by head_dim^-0.5 = 24^-0.5
    qkv = sd["audio_encoder.base.htsat.layers.0.blocks.0.attn.qkv.weight"].clone()
    qkv[:96] *= 24 ** -0.5
    assert (planes("s0.b0.qkv_w").view_as(qkv) - qkv).abs().max() < 1e-5
    # rotate-half partners are adjacent after the q/k row permutation; v rows untouched
    q = sd["caption_decoder.lm.model.layers.0.self_attn.q_proj.weight"]
    packed = planes("lm.l0.qkv_w").view(960, 576)
    assert torch.allclose(packed[0], q[0], atol=1e-5) and torch.allclose(packed[1], q[32], atol=1e-5)
    assert torch.allclose(packed[64 + 3], q[64 + 33], atol=1e-5)
    v = sd["caption_decoder.lm.model.layers.0.self_attn.v_proj.weight"]
    assert torch.allclose(packed[768:], v, atol=1e-5)
    gate = sd["caption_decoder.lm.model.layers.0.mlp.gate_proj.weight"]
    up = sd["caption_decoder.lm.model.layers.0.mlp.up_proj.weight"]
    gu = planes("lm.l0.gu_w").view(3072, 576)
    assert torch.allclose(gu[10], gate[5], atol=1e-5) and torch.allclose(gu[11], up[5], atol=1e-5)
    # folded BatchNorm
    p = "audio_encoder.base.htsat.bn0."
    x = torch.linspace(-60, 10, 64)
    want = (x - sd[p + "running_mean"]) / torch.sqrt(sd[p + "running_var"] + 1e-5) * sd[p + "weight"] + sd[p + "bias"]
    assert torch.allclose(x * f32("fe.bn_scale") + f32("fe.bn_shift"), want, atol=1e-5)
    # mel filter extents cover every non-zero weight
    melW = sd["audio_encoder.base.htsat.logmel_extractor.melW"]
    o, b = table["fe.mel_lo"]; lo = arena[o:o + b].view(torch.int32)
    o, b = table["fe.mel_hi"]; hi = arena[o:o + b].view(torch.int32)
    for m in range(64):
        nz = torch.nonzero(melW[:, m]).flatten()
        assert lo[m] <= nz.min() and hi[m] > nz.max()
    # rope table equals the oracle's
    from oracle import restated as R
    cos, sin = R.rope_tables(S.MAX_POSITIONS)
    assert torch.equal(f32("lm.rope_cos").view(-1, 32), cos[:, :32]) and torch.equal(f32("lm.rope_sin").view(-1, 32), sin[:, :32])


def test_packer_rejects_bad_checkpoints(sd, lib):
    entries = native.weight_entries(lib)
    bad = dict(sd)
    k = "audio_encoder.base.htsat.spectrogram_extractor.stft.conv_imag.weight"
    bad[k] = sd[k] * 1.01
    with pytest.raises(RuntimeError, match="windowed DFT"):
        weights.pack(bad, entries, lib.mb_weights_size())
    bad = dict(sd)
    del bad["audio_encoder.base.c2l.bias"]
    with pytest.raises(RuntimeError, match="schema"):
        weights.pack(bad, entries, lib.mb_weights_size())
    wrapped = {"module." + k: v for k, v in sd.items()}
    assert set(weights.strip_module_prefix(wrapped)) == set(sd)          # wrapper.py:77-82


def test_synthetic_checkpoint_is_reproducible(sd):
    again = synth.synthetic_state_dict()
    for k in ("audio_encoder.projection.linear1.weight", "caption_decoder.lm.model.layers.29.mlp.down_proj.weight"):
        assert torch.equal(sd[k], again[k])
    import hashlib
    h = hashlib.sha256(sd["caption_decoder.lm.model.norm.weight"].numpy().tobytes()).hexdigest()
    assert h == hashlib.sha256(again["caption_decoder.lm.model.norm.weight"].numpy().tobytes()).hexdigest()
    assert S.count_parameters() == S.TOTAL_PARAMS
    assert S.PREFIX_LEN == 3 * 129 + 2 and S.N_FRAMES == 1001


def test_shift_mask_closed_form():
    m = dsp.shifted_window_mask(16)
    assert m.shape == (4, 64, 64) and set(np.unique(m)) == {-100.0, 0.0}
    assert (m[0] == 0).all()                                           # top-left window is not cut by the shift


def test_tokenizer_stand_in_and_wrapper_errors():
    from mellow_b200.tokenizer import ByteStandInTokenizer, tokenize_prompts
    tok = ByteStandInTokenizer()
    ids = tokenize_prompts(tok, ["hi!", "what?"], 129)
    assert ids.shape == (2, 129) and ids.dtype == torch.int64
    assert ids[0, 3:].eq(17).all() and tok.encode("<|endoftext|>")[0] == 0
    assert tok.decode([256 + ord("o"), 256 + ord("k"), 0, 300]).split("<|endoftext|>")[0] == "ok"
    from mellow_b200 import MellowWrapper, MellowNativeError
    with pytest.raises(ValueError):
        MellowWrapper(config="v0", model="v9", device=0)
    if not torch.cuda.is_available():
        with pytest.raises(MellowNativeError):
            MellowWrapper(config="v0", model="v0", device="cpu", use_cuda=False, checkpoint="synthetic")


def test_tokenizer_never_falls_back_silently(monkeypatch):
    """without a local tokenizer and without the hub, loading must RAISE (the reference raises too);
    the stand-in is used only when asked for."""
    from mellow_b200.tokenizer import ByteStandInTokenizer, load_tokenizer
    monkeypatch.delenv("MELLOW_TOKENIZER", raising=False)
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    with pytest.raises(RuntimeError, match="cannot load the tokenizer"):
        load_tokenizer("HuggingFaceTB/SmolLM2-135M-not-cached-anywhere")
    assert isinstance(load_tokenizer("HuggingFaceTB/SmolLM2-135M", local="stand-in"), ByteStandInTokenizer)
    monkeypatch.setenv("MELLOW_TOKENIZER", "stand-in")
    assert isinstance(load_tokenizer("HuggingFaceTB/SmolLM2-135M"), ByteStandInTokenizer)


def test_resampled_length_rounds_like_torchaudio():
    """torchaudio ceils the float32-rounded quotient; 1 % of the 44.1 kHz lengths differ from math.ceil."""
    from mellow_b200.audio_io import resampled_length
    assert resampled_length(300044, 44100, 32000) == 217719 == resampled_length(300044, 441, 320)
    for n in list(range(300000, 300400)) + [403604, 445940, 1, 2, 441, 440999]:
        want = int(torch.ceil(torch.as_tensor(320 * n / 441)).long())
        assert resampled_length(n, 44100, 32000) == want, n
    assert resampled_length(403604, 44100, 32000) == 292865 and resampled_length(445940, 44100, 32000) == 323585


def test_global_stop_rule_of_split_runs():
    """wrapper.py:247-249 evaluated after the fact (runs split into passes / ranks): columns up to and including the
    first step at which every row has emitted the stop id."""
    from mellow_b200.wrapper import MellowWrapper
    t = torch.tensor([[5, 0, 7, 7, 7], [6, 6, 6, 0, 9], [0, 1, 2, 3, 4]])
    assert MellowWrapper._global_stop(t, 0) == 4
    assert MellowWrapper._global_stop(t, 8) == 5                      # never reached: max_len columns
    assert MellowWrapper._global_stop(t[:1], 0) == 2


def test_read_audio_info_and_errors(tmp_path):
    import wave as wavmod
    from mellow_b200.audio_io import read_audio
    p = str(tmp_path / "s.wav")
    with wavmod.open(p, "wb") as f:
        f.setnchannels(2); f.setsampwidth(2); f.setframerate(48000)
        f.writeframes(np.arange(2 * 100, dtype="<i2").tobytes())
    assert read_audio(p, info_only=True) == (2, 100, 48000)
    x, sr = read_audio(p)
    assert x.shape == (2, 100) and sr == 48000 and x[1, 0] == 1 / 32768.0
    bad = tmp_path / "x.flac"
    bad.write_bytes(b"fLaC" + bytes(64))
    with pytest.raises(RuntimeError, match="cannot decode"):
        read_audio(str(bad))


def test_create_fails_loudly_without_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = lib.mb_create(0, 1, 8, 0)
    assert not h
    assert b"no CUDA device" in lib.mb_last_error(None)


def test_resample_filter_bank_equals_torchaudio():
    """the polyphase filter bank handed to the GPU resampler is torchaudio's, bit for bit."""
    import torchaudio.transforms as T
    from mellow_b200.audio_io import resampled_length, sinc_resample_kernel
    for sr_in in (44100, 48000, 22050, 16000, 8000):
        kern, width, orig, new = sinc_resample_kernel(sr_in, 32000)
        ref = T.Resample(sr_in, 32000)
        assert width == ref.width and kern.shape[1] == 2 * width + orig
        assert torch.equal(kern, ref.kernel[:, 0, :])
        x = torch.zeros(1, 4567)
        assert ref(x).shape[1] == resampled_length(4567, orig, new)


def test_plan_fit_draws_like_the_reference():
    import random
    from mellow_b200.audio_io import plan_fit
    assert plan_fit(1000, 320000, random.Random(1)) == 0                   # short clip: tiled, no draw
    rng = random.Random(7)
    want = random.Random(7).randrange(323585 - 320000)
    assert plan_fit(323585, 320000, rng) == want                          # long clip: one randrange draw


def test_local_hf_tokenizer_plumbing(tmp_path):
    """with a local tokenizer directory (tokenizer= / $MELLOW_TOKENIZER) the wrapper uses a real
    Hugging Face fast tokenizer the way the reference does (wrapper.py:84-85,186-190,254): pad token '!', right padding /
    truncation to 129 ids, no BOS / EOS added, stop id = first id of '<|endoftext|>', text cut at the stop token.  The
    SmolLM2 files are not available offline, so a small byte-level BPE tokenizer with the same special token is built."""
    tokenizers = pytest.importorskip("tokenizers")
    from tokenizers import Tokenizer, models, pre_tokenizers, decoders, trainers
    from mellow_b200.tokenizer import load_tokenizer, tokenize_prompts
    tk = Tokenizer(models.BPE())
    tk.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tk.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=400, special_tokens=["<|endoftext|>"],
                                  initial_alphabet=pre_tokenizers.ByteLevel.alphabet())
    tk.train_from_iterator(["what is the difference between the two audios?", "describe the audio in detail!",
                            "is there a dog barking? yes! no!"] * 4, trainer)
    tk.save(str(tmp_path / "tokenizer.json"))
    (tmp_path / "tokenizer_config.json").write_text(
        '{"tokenizer_class": "PreTrainedTokenizerFast", "eos_token": "<|endoftext|>", "bos_token": "<|endoftext|>", '
        '"unk_token": "<|endoftext|>", "model_max_length": 8192}')
    tok = load_tokenizer("HuggingFaceTB/SmolLM2-135M", local=str(tmp_path))
    assert not getattr(tok, "is_stand_in", False)
    assert tok.pad_token == "!" and tok.pad_token_id == tok.encode("!")[0]
    stop_id = tok.encode("<|endoftext|>")[0]
    assert stop_id == 0                                                # first special token, like SmolLM2
    prompts = ["what is the difference between the two audios?", "yes! " * 200]
    ids = tokenize_prompts(tok, prompts, 129)
    assert ids.shape == (2, 129) and ids.dtype == torch.int64
    n0 = len(tok.encode(prompts[0]))
    assert ids[0, :n0].tolist() == tok.encode(prompts[0])              # no BOS / EOS added
    assert ids[0, n0:].eq(tok.pad_token_id).all()                      # right padding with '!'
    assert ids[1].tolist() == tok.encode(prompts[1])[:129]             # truncation
    # detokenisation rule of wrapper.py:251-254
    row = tok.encode("a dog barking") + [stop_id] + tok.encode("garbage after the stop token")
    assert tok.decode(row).split("<|endoftext|>")[0] == "a dog barking"
    with pytest.raises(Exception):
        load_tokenizer("HuggingFaceTB/SmolLM2-135M", local=str(tmp_path / "missing"))
