"""CPU, build container only: the oracle restatement against the UNMODIFIED reference classes imported from
/root/reference (skipped where that tree does not exist, e.g. the GPU box)."""
import pytest
import torch

from oracle import reference_model as RM
from oracle import restated as R

pytestmark = pytest.mark.skipif(not RM.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref_model(sd):
    return RM.build_reference_model(sd)


def test_schema_and_param_count(sd, ref_model):
    from mellow_b200 import schema as S
    ref = ref_model.state_dict()
    assert list(ref.keys()) == list(S.checkpoint_schema().keys())
    for k, (shape, dtype) in S.checkpoint_schema().items():
        assert tuple(ref[k].shape) == tuple(shape) and str(ref[k].dtype) == "torch." + dtype
    assert sum(p.numel() for p in ref_model.parameters()) == S.TOTAL_PARAMS == S.count_parameters()


def test_prefix_matches_reference_on_other_inputs(sd, ref_model):
    from mellow_b200 import synth
    wave = synth.synthetic_waveforms(2, seed=99)
    ids = synth.synthetic_prompt_ids(1, n_real=20, seed=99)
    with torch.no_grad():
        want, _, _ = ref_model.generate_prefix_inference({"audio1": wave[:1], "audio2": wave[1:], "input": {"input_ids": ids}})
        got = R.build_prefix(sd, R.encode_clips(sd, wave[:1]), R.encode_clips(sd, wave[1:]), ids)
        assert (want - got).abs().max() < 5e-5
        ref_logits = ref_model.caption_decoder.lm(inputs_embeds=want).logits[:, -1]
        my_logits = R.last_logits(sd, R.llama_hidden(sd, got))
        assert (ref_logits - my_logits).abs().max() < 2e-4


def test_audio_prep_on_reference_wavs(sd):
    """wrapper.py:141-168 on resource/1.wav (tiled) and 2.wav (cropped): SURVEY.md Appendix B."""
    import os
    import random
    from mellow_b200.audio_io import load_audio_into_tensor, read_wav
    res = os.path.join(RM.REFERENCE_ROOT, "resource")
    a1, sr1 = read_wav(os.path.join(res, "1.wav"))
    assert sr1 == 44100 and a1.shape == (1, 403604)
    random.seed(0)
    x1 = load_audio_into_tensor(os.path.join(res, "1.wav"), 10, 32000, True, random)
    assert x1.shape == (320000,)
    assert torch.equal(x1[:27135], x1[292865:320000])                   # 292865 resampled samples, tiled
    random.seed(0)
    start = random.Random(0).randrange(323585 - 320000)
    x2 = load_audio_into_tensor(os.path.join(res, "2.wav"), 10, 32000, True, random)
    import torchaudio.transforms as T
    a2, sr2 = read_wav(os.path.join(res, "2.wav"))
    full = T.Resample(sr2, 32000)(a2).reshape(-1)
    assert full.shape[0] == 323585
    assert torch.equal(x2, full[start:start + 320000])
