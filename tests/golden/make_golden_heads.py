"""Generates tests/golden/ref_heads_synth1234.npz: the encoder output dict the reference returns as od1 / od2 from
``Mellow.generate_prefix_inference`` (mellow/model/mellow.py:100-108), produced by the
UNMODIFIED reference classes (oracle/reference_model.py) on the seeded synthetic checkpoint and inputs.  Only
runnable where /root/reference exists (the build container):

    python tests/golden/make_golden_heads.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mellow_b200 import synth  # noqa: E402
from oracle.reference_model import build_reference_model  # noqa: E402

B = 2


def main():
    torch.manual_seed(0)
    sd = synth.synthetic_state_dict()
    model = build_reference_model(sd)
    wave = synth.synthetic_waveforms(2 * B)
    ids = synth.synthetic_prompt_ids(B)
    g = {}
    with torch.no_grad():
        d = {"audio1": wave[:B], "audio2": wave[B:], "input": {"input_ids": ids}}
        _, od1, od2 = model.generate_prefix_inference(d)
        for name, od in (("od1", od1), ("od2", od2)):
            fw = od["framewise_output"]                                   # (B,1024,527): 32 unique rows, each 32x
            assert fw.shape == (B, 1024, 527) and torch.equal(fw[:, 0::32], fw[:, 31::32])
            emb = od["embedding"]
            assert emb.shape == (B, 1025, 768) and torch.equal(emb[:, 1::32], emb[:, 32::32])
            g[name + "_clipwise"] = od["clipwise_output"].numpy()
            g[name + "_framewise_rows"] = fw[:, 0::32].numpy()
            g[name + "_latent"] = od["latent_output"].numpy()
            g[name + "_embedding_rows"] = torch.cat([emb[:, :1], emb[:, 1::32]], dim=1).numpy()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_heads_synth1234.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")
    print("clipwise range", float(g["od1_clipwise"].min()), float(g["od1_clipwise"].max()))


if __name__ == "__main__":
    main()
