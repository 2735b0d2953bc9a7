"""GPU parity at the lengths BASELINE.json names (max_len = 300, ctx 389..688; configs[0] on the reference's own wavs),
against fixtures produced by the UNMODIFIED reference classes under the reference's cache-less loop
(tests/golden/make_golden_long.py).

What "identical greedy ids" can mean over 300 steps: the reference's own decision margin (top-1 minus top-2 logit)
drops to 2e-4 on some steps of these synthetic checkpoints, far below what fp32 re-association alone moves a logit
(the KV-cached loop here vs the reference's cache-less loop; a different BLAS thread count on the CPU does the same).
The gate is therefore stated against the tolerance: LOGIT_TOL = 1e-2 (the north-star "per-step logits within a stated
tolerance"), and
  * teacher-forced on the reference's ids, EVERY step's top-8 and probe logits are within LOGIT_TOL and the model's
    own argmax equals the reference id wherever the reference margin is >= MARGIN_GATE = 2 * LOGIT_TOL;
  * free-running, every row's ids are identical up to its first step whose reference margin is below MARGIN_GATE
    (all 300 steps when there is none).
"""
import importlib.util
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
LOGIT_TOL = 1e-2
MARGIN_GATE = 2 * LOGIT_TOL


def _maker():
    spec = importlib.util.spec_from_file_location("make_golden_long", os.path.join(GOLD, "make_golden_long.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                     # imports oracle.reference_model lazily-safe: no reference access at import
    return mod


def _golden(name):
    path = os.path.join(GOLD, f"ref_{name}.npz")
    if not os.path.isfile(path):
        pytest.skip(f"{path} not generated")
    return dict(np.load(path))


@pytest.fixture(scope="module")
def arenas(engine):
    """checkpoint seed -> device arena (seed 1234 is the session engine's)."""
    from mellow_b200 import synth
    cache = {1234: engine.arena}

    def get(seed):
        if seed not in cache:
            cache[seed] = engine.pack_arena(synth.synthetic_state_dict(seed)).to(engine.device)
        return cache[seed]
    return get


def _first_thin_step(margins_row):
    thin = np.nonzero(margins_row < MARGIN_GATE)[0]
    return int(thin[0]) if thin.size else margins_row.shape[0]


@pytest.mark.parametrize("policy", ["split", "split24"])
@pytest.mark.parametrize("name", ["long1234", "rows1234", "rows77"])
def test_300_steps_against_the_reference_loop(name, policy, arenas):
    from mellow_b200.engine import Engine
    g = _golden(name)
    seed, w1, w2, ids, steps = _maker().set_inputs(name)
    assert np.array_equal(ids.numpy(), g["input_ids"])
    ref = torch.from_numpy(g["tokens"]).to(torch.int32)                      # (B, 300)
    B = ref.shape[0]
    assert ref.shape[1] == steps == 300
    margins = (g["top8_vals"][..., 0] - g["top8_vals"][..., 1]).T            # (B, steps)
    eng = Engine(None, device=0, max_batch=B, max_new_tokens=steps, policy=policy, arena=arenas(seed))
    try:
        # free-running greedy loop (CUDA graph, KV cache) vs the reference's cache-less loop
        free = eng.generate(w1, w2, ids, steps).cpu()
        assert free.shape == (B, steps)
        identical_rows = 0
        for b in range(B):
            upto = _first_thin_step(margins[b])
            assert torch.equal(free[b, :upto], ref[b, :upto]), (
                f"{name}/{policy} row {b}: ids differ at step {int((free[b, :upto] != ref[b, :upto]).nonzero()[0])} "
                f"before the first thin-margin step {upto}")
            identical_rows += int(torch.equal(free[b], ref[b]))
        # teacher-forced on the reference ids: every step's logits, to ctx 688
        eng.encode(w1, w2)
        eng.prefix(ids)
        eng.prefill(B, want_logits=False)
        own, dump = eng.decode(B, steps, dump_logits=True, forced_tokens=ref)
        own, dump = own.cpu(), dump.cpu()                                     # (B, steps), (steps, B, V)
        assert torch.isfinite(dump).all()
        top_ids = torch.from_numpy(g["top8_ids"])                             # (steps, B, 8)
        got_top = torch.gather(dump, 2, top_ids)
        err_top = (got_top - torch.from_numpy(g["top8_vals"])).abs().max().item()
        got_probe = dump[:, :, torch.from_numpy(g["probe_ids"])]
        err_probe = (got_probe - torch.from_numpy(g["probe_logits"])).abs().max().item()
        assert err_top < LOGIT_TOL and err_probe < LOGIT_TOL, f"{name}/{policy}: top-8 {err_top:.2e}, probes {err_probe:.2e}"
        decided = torch.from_numpy(margins >= MARGIN_GATE)
        assert torch.equal(own[decided], ref[decided]), f"{name}/{policy}: argmax differs on a step with margin >= {MARGIN_GATE}"
        agree = (own == ref).float().mean().item()
        print(f"{name}/{policy}: teacher-forced max |dlogit| top-8 {err_top:.2e} probes {err_probe:.2e}; argmax agrees on "
              f"{agree:.4f} of {B * steps} steps ({int((~decided).sum())} below the margin gate); free-running rows "
              f"identical over all 300 steps: {identical_rows}/{B}")
    finally:
        eng.close()


def test_batch_128_runs_300_steps_like_the_reference(engine):
    """The bench configuration itself (B=128, max_len 300, default policy): 64 copies of the two long1234 rows."""
    from mellow_b200.engine import Engine
    g = _golden("long1234")
    _, w1, w2, ids, steps = _maker().set_inputs("long1234")
    ref = torch.from_numpy(g["tokens"]).to(torch.int32)
    margins = (g["top8_vals"][..., 0] - g["top8_vals"][..., 1]).T
    eng = Engine(None, device=0, max_batch=128, max_new_tokens=steps, policy="split24", arena=engine.arena)
    try:
        toks = eng.generate(w1.repeat(64, 1), w2.repeat(64, 1), ids.repeat(64, 1), steps).cpu()
        assert toks.shape == (128, steps)
        for b in range(2):
            rows = toks[b::2]
            assert (rows == rows[:1]).all(), "copies of one pair must decode identically whatever their row index"
            upto = _first_thin_step(margins[b])
            assert torch.equal(rows[0, :upto], ref[b, :upto])
    finally:
        eng.close()


def test_finished_rows_hand_their_attention_ctas_to_the_live_rows(engine):
    """SURVEY 8 row f3 (option share_keys): rows that emitted the stop token leave the batch on the device; their
    decode-attention CTAs then take a share of the keys of the rows still decoding (2, then 4 CTAs per row and kv head, then
    idle slots).  The live rows must keep producing the reference's logits / ids while 64, 96 and 112 of the 128 rows
    finish (teacher forcing: the forced stop token sets the row's done flag)."""
    from mellow_b200.engine import Engine
    g = _golden("long1234")
    _, w1, w2, ids, steps = _maker().set_inputs("long1234")
    ref = torch.from_numpy(g["tokens"]).to(torch.int32)                      # (2, 300)
    margins = (g["top8_vals"][..., 0] - g["top8_vals"][..., 1]).T            # (2, steps)
    B, live = 128, 16
    forced = ref.repeat(64, 1).contiguous()                                  # row r is a copy of pair r % 2
    forced[64:, 2] = 0                                                       # 64 rows stop: 2 CTAs per live row and kv head
    forced[32:64, 4] = 0                                                     # 96 stopped: 4 CTAs
    forced[16:32, 8] = 0                                                     # 112 stopped: 4 CTAs, 64 idle slots
    eng = Engine(None, device=0, max_batch=B, max_new_tokens=steps, policy="split24", arena=engine.arena)
    try:
        eng.encode(w1.repeat(64, 1), w2.repeat(64, 1))

        def run(n, **kw):
            eng.prefix(ids.repeat(64, 1))
            eng.prefill(B, want_logits=False)
            return eng.decode(B, n, forced_tokens=forced[:, :n].contiguous().cuda(), **kw)

        # (i) logits of the live rows over the steps that cross the three hand-overs (individual launches)
        n = 16
        own, dump = run(n, dump_logits=True)
        dump = dump[:, :live].cpu()                                          # (n, live, V)
        assert torch.isfinite(dump).all()
        top_ids = torch.from_numpy(g["top8_ids"][:n]).repeat(1, live // 2, 1)   # (n, live, 8): row r -> pair r % 2
        top_vals = torch.from_numpy(g["top8_vals"][:n]).repeat(1, live // 2, 1)
        err = (torch.gather(dump, 2, top_ids) - top_vals).abs().max().item()
        assert err < LOGIT_TOL, f"live rows' top-8 logits off by {err:.2e} while finished rows hand over their CTAs"
        # (ii) 300 steps under the CUDA graph: ids of the live rows, against the reference and against share_keys = 0
        shared = run(steps).cpu()
        eng.set_option("share_keys", 0)
        plain = run(steps).cpu()
        decided = torch.from_numpy(margins >= MARGIN_GATE)                   # (2, steps)
        for r in range(live):
            d = decided[r % 2]
            assert torch.equal(shared[r][d], ref[r % 2][d]), f"row {r}: ids differ from the reference on a decided step"
            assert torch.equal(shared[r][d], plain[r][d]), f"row {r}: share_keys changes a decided id"
        print(f"share_keys: live-row top-8 logits within {err:.2e}; ids identical to share_keys=0 on "
              f"{(shared[:live] == plain[:live]).float().mean().item():.4f} of the steps")
    finally:
        eng.close()


def test_config0_reference_wavs_through_the_wrapper():
    """BASELINE.json configs[0]: resource/1.wav + 2.wav, random.seed(0), 30 greedy steps through MellowWrapper.generate()
    (GPU resampler, tile / crop, stand-in tokenizer) against ids produced by the reference classes from audio prepared
    exactly like wrapper.py:141-168."""
    from mellow_b200 import MellowWrapper
    from mellow_b200.tokenizer import ByteStandInTokenizer
    g = _golden("config0")
    res = os.path.join(GOLD, "resource")
    prompt = _maker().CONFIG0_PROMPT
    mw = MellowWrapper(config="v0", model="v0", device=0, use_cuda=True, checkpoint="synthetic")
    random.seed(0)
    out = mw.generate(examples=[[os.path.join(res, "1.wav"), os.path.join(res, "2.wav"), prompt]], max_len=30, top_p=0.8,
                      temperature=1.0)
    want = ByteStandInTokenizer().decode(g["tokens"][0].tolist()).split("<|endoftext|>")[0]
    assert out == [want]
    # the audio the GPU ingest produced is the reference's (tile for 1.wav, the seeded crop for 2.wav)
    random.seed(0)
    a1 = mw.preprocess_audio([os.path.join(res, "1.wav")], True)
    a2 = mw.preprocess_audio([os.path.join(res, "2.wav")], True)
    assert (a1[0, :4096].cpu() - torch.from_numpy(g["audio1_head"])).abs().max() < 5e-6
    assert (a2[0, :4096].cpu() - torch.from_numpy(g["audio2_head"])).abs().max() < 5e-6
    mw.model.close()


def test_reference_inner_seams(engine, sd):
    """wrapper.py:217,237 call `model.caption_decoder.lm(inputs_embeds=...)` and `lm.model.embed_tokens`; the cache-less
    forward over prefix + t embeddings must give the reference's step-t logits (golden: t = 11) and the KV-cached loop."""
    from oracle import restated as R
    g = dict(np.load(os.path.join(GOLD, "ref_synth1234.npz")))
    _, w1, w2, ids, _ = _maker().set_inputs("long1234")
    prefix, od1, od2 = engine.generate_prefix_inference({"audio1": w1, "audio2": w2, "input": {"input_ids": ids}}, heads=False)
    lm = engine.caption_decoder.lm
    toks = torch.from_numpy(g["tokens"])                                      # (2, 12)
    emb = lm.model.embed_tokens(toks[:, :11].cuda())
    assert torch.equal(emb.cpu(), sd["caption_decoder.lm.model.embed_tokens.weight"][toks[:, :11]])
    seq = torch.cat([prefix, emb], dim=1)                                     # (2, 400, 576)
    logits = lm(inputs_embeds=seq).logits[:, -1, :].cpu()
    top_ids = torch.from_numpy(g["top8_ids"][11])
    err = (torch.gather(logits, 1, top_ids) - torch.from_numpy(g["top8_vals"][11])).abs().max().item()
    assert err < LOGIT_TOL, err
    assert logits.argmax(-1).tolist() == toks[:, 11].tolist()
    with torch.no_grad():
        want0 = R.last_logits(sd, R.llama_hidden(sd, prefix.cpu()))
    got0 = lm(inputs_embeds=prefix).logits[:, -1, :].cpu()
    assert (got0 - want0).abs().max() < LOGIT_TOL


def _write_wavs(tmp_path, n, seed=321):
    """n mono 16-bit wav files from the seeded synthetic waveforms, mixed rates / lengths (tile and crop branches)."""
    import wave as wavmod
    from mellow_b200 import synth
    w = synth.synthetic_waveforms(n, seed=seed)
    paths = []
    for i in range(n):
        sr = [32000, 44100, 22050][i % 3]
        seconds = [10.0, 10.7, 4.2][i % 3]
        x = torch.nn.functional.interpolate(w[i][None, None, :], size=int(sr * seconds), mode="linear")[0, 0]
        p = str(tmp_path / f"clip{i}.wav")
        with wavmod.open(p, "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(sr)
            f.writeframes((x.numpy() * 32767.0).round().astype("<i2").tobytes())
        paths.append(p)
    return paths


def test_wrapper_passes_and_custom_stop_token(tmp_path):
    """MellowWrapper.generate() over more examples than one engine pass holds (capacity 2 -> three passes for 5 examples)
    returns what a single pass returns, also with a stop token other than '<|endoftext|>' -- the reference then returns
    every row up to the GLOBAL stop step (wrapper.py:247-254), which the wrapper applies across passes."""
    from mellow_b200 import MellowWrapper
    paths = _write_wavs(tmp_path, 10)
    examples = [[paths[i], paths[5 + i], f"what differs? ({i})"] for i in range(5)]
    one = MellowWrapper(config="v0", model="v0", device=0, use_cuda=True, checkpoint="synthetic")
    many = MellowWrapper(config="v0", model="v0", device=0, use_cuda=True, checkpoint="synthetic", max_batch=2)
    try:
        for stop in ("<|endoftext|>", "!"):
            random.seed(11)
            a = one.generate(examples=examples, max_len=10, top_p=0.8, temperature=1.0, stop_token=stop)
            random.seed(11)
            b = many.generate(examples=examples, max_len=10, top_p=0.8, temperature=1.0, stop_token=stop)
            assert a == b and len(a) == 5 and len(set(a)) > 1, stop
        assert many.model.max_batch == 2 and one.model.max_batch >= 5
        # a longer request than the handle was sized for re-creates it (the reference accepts any max_len)
        random.seed(11)
        c = one.generate(examples=examples[:1], max_len=320, top_p=0.8, temperature=1.0)
        assert one.model.max_new_tokens >= 320 and len(c) == 1
    finally:
        one.model.close()
        many.model.close()


def test_checkpoint_file_and_local_hf_tokenizer_end_to_end(tmp_path, sd):
    """SURVEY section 8 row f2 as far as it goes offline: a checkpoint FILE in the reference's format (torch.save of the
    flat state_dict, with the DataParallel 'module.' prefix the reference strips at wrapper.py:77-82) and a local Hugging
    Face fast tokenizer directory; the text must be what the CPU oracle's ids decode to with that tokenizer."""
    tokenizers = pytest.importorskip("tokenizers")
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers
    from mellow_b200 import MellowWrapper
    from mellow_b200.audio_io import load_audio_into_tensor
    from oracle import restated as R
    tokdir = tmp_path / "tok"
    tokdir.mkdir()
    tk = Tokenizer(models.BPE())
    tk.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tk.decoder = decoders.ByteLevel()
    trainer = trainers.BpeTrainer(vocab_size=400, special_tokens=["<|endoftext|>"],
                                  initial_alphabet=pre_tokenizers.ByteLevel.alphabet())
    tk.train_from_iterator(["what is the difference between the two audios?", "describe the audio in detail!"] * 4, trainer)
    tk.save(str(tokdir / "tokenizer.json"))
    (tokdir / "tokenizer_config.json").write_text(
        '{"tokenizer_class": "PreTrainedTokenizerFast", "eos_token": "<|endoftext|>", "bos_token": "<|endoftext|>", '
        '"unk_token": "<|endoftext|>", "model_max_length": 8192}')
    ckpt = str(tmp_path / "v0.ckpt")
    torch.save({"module." + k: v for k, v in sd.items()}, ckpt)
    paths = _write_wavs(tmp_path, 2, seed=77)
    prompt = "what is the difference between the two audios?"
    mw = MellowWrapper(config="v0", model="v0", device=0, use_cuda=True, checkpoint=ckpt, tokenizer=str(tokdir))
    try:
        random.seed(3)
        out = mw.generate(examples=[[paths[0], paths[1], prompt]], max_len=8, top_p=0.8, temperature=1.0)
        ids = mw.preprocess_text([prompt])["input_ids"]
        assert ids.shape == (1, 129) and ids[0, -1] == mw.tokenizer.pad_token_id
        random.seed(3)
        a1 = load_audio_into_tensor(paths[0], 10, 32000, True, random)[None]
        a2 = load_audio_into_tensor(paths[1], 10, 32000, True, random)[None]
        with torch.no_grad():
            want = R.generate_from_wave(sd, a1, a2, ids, 8)
        assert out == [mw.tokenizer.decode(want[0].tolist()).split("<|endoftext|>")[0]]
    finally:
        mw.model.close()
