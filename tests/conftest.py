import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_synth1234.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sd():
    from mellow_b200 import synth
    return synth.synthetic_state_dict()


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="session")
def golden_heads():
    """od1 / od2 of the reference's generate_prefix_inference (tests/golden/make_golden_heads.py)."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ref_heads_synth1234.npz")))


@pytest.fixture(scope="session")
def inputs():
    """The seeded synthetic batch the golden fixture was generated from (B=2 pairs)."""
    from mellow_b200 import synth
    wave = synth.synthetic_waveforms(4)
    return {"wave1": wave[:2], "wave2": wave[2:], "ids": synth.synthetic_prompt_ids(2)}


@pytest.fixture(scope="session")
def engine(sd):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mellow_b200.engine import Engine
    eng = Engine(sd, device=0, max_batch=4, max_new_tokens=32, policy="split")
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engine24(sd, engine):
    """policy split24: split operands, KV cache stored at 24 bits."""
    from mellow_b200.engine import Engine
    eng = Engine(None, device=0, max_batch=7, max_new_tokens=32, policy="split24", arena=engine.arena)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engine_fast(sd, engine):
    from mellow_b200.engine import Engine
    eng = Engine(None, device=0, max_batch=2, max_new_tokens=32, policy="fast", arena=engine.arena)
    yield eng
    eng.close()
