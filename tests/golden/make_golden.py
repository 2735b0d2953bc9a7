"""Generates tests/golden/ref_synth1234.npz by running the UNMODIFIED reference model classes
(/root/reference/mellow/model/*.py, imported through oracle/reference_model.py) on the seeded synthetic
checkpoint and inputs.  Only runnable where /root/reference exists (the build container):

    python tests/golden/make_golden.py

The fixture pins (a) the standalone oracle restatement (tests/test_oracle_golden.py, CPU) and (b) the CUDA path
(tests/test_gpu_parity.py, GPU) to outputs of the reference's own code.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mellow_b200 import synth  # noqa: E402
from oracle.reference_model import build_reference_model, reference_generate_ids  # noqa: E402

B, STEPS = 2, 12
TOK_ROWS = {"patch": [0, 2047, 4095], "stage0": [0, 511, 1023], "stage1": [0, 100, 255], "stage2": [0, 31, 63],
            "stage3": [0, 31, 63]}
FRAME_ROWS = [0, 1, 2, 500, 998, 999, 1000]


def main():
    torch.manual_seed(0)
    sd = synth.synthetic_state_dict()
    model = build_reference_model(sd)
    wave = synth.synthetic_waveforms(2 * B)
    ids = synth.synthetic_prompt_ids(B)
    g = {}
    with torch.no_grad():
        ht = model.audio_encoder.base.htsat
        w1 = wave[:B]
        spec = ht.spectrogram_extractor(w1)
        lm = ht.logmel_extractor(spec)
        g["logmel_rows"] = lm[:, 0, FRAME_ROWS, :].numpy()
        x = ht.bn0(lm.transpose(1, 3)).transpose(1, 3)
        g["bn_rows"] = x[:, 0, FRAME_ROWS, :].numpy()
        img = ht.reshape_wav2img(x)
        tok = ht.patch_embed(img)
        g["patch"] = tok[:, TOK_ROWS["patch"], :].numpy()
        for i, layer in enumerate(ht.layers):
            tok, _ = layer(tok)
            g[f"stage{i}"] = tok[:, TOK_ROWS[f"stage{i}"], :].numpy()
        proj1, _, od1 = model.audio_encoder(w1)
        g["latent"] = od1["latent_output"].numpy()
        g["frames"] = od1["embedding"][:, 1::32, :].numpy()          # 32 unique c2l rows
        g["rows33"] = torch.cat([proj1[:, :1], proj1[:, 1::32]], dim=1).numpy()
        d = {"audio1": wave[:B], "audio2": wave[B:], "input": {"input_ids": ids}}
        prefix, _, _ = model.generate_prefix_inference(d)
        g["prefix_rows"] = prefix[:, [0, 1, 4, 5, 128, 129, 130, 131, 258, 259, 260, 323, 324, 388], :].numpy()
        toks, logits = reference_generate_ids(model, prefix, STEPS, top_p=0.8, temperature=1.0, dump_logits=True)
        logits = torch.stack(logits, 0)                              # (steps, B, V)
        g["tokens"] = toks.numpy().astype(np.int64)
        top = logits.topk(8, dim=-1)
        g["top8_ids"] = top.indices.numpy().astype(np.int64)
        g["top8_vals"] = top.values.numpy()
        probe = np.random.Generator(np.random.PCG64(7)).integers(0, 49152, size=64)
        g["probe_ids"] = probe.astype(np.int64)
        g["probe_logits"] = logits[:, :, torch.from_numpy(probe)].numpy()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_synth1234.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")
    print("tokens", g["tokens"].tolist())
    print("top1-top2 margins", (g["top8_vals"][..., 0] - g["top8_vals"][..., 1]).round(3).tolist())


if __name__ == "__main__":
    main()
