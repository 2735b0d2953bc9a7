"""CPU, gloo, world size 2: the N > 1 path of ``MellowWrapper.generate()`` -- contiguous example slices per rank, the
reference's ``random.randrange`` draws made for EVERY clip on every rank (so a sharded run crops where the single
process does), passes of the engine's capacity, host gather in example order, the global stop rule for custom stop
tokens.  The native engine is replaced by a deterministic stand-in (no GPU here); the 2-GPU run of the real engine is
tests/test_gpu_multi.py."""
import os
import random
import socket
import wave as wavmod

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mellow_b200 import schema as S
from mellow_b200.audio_io import plan_fit, resampled_length
from mellow_b200.tokenizer import ByteStandInTokenizer
from mellow_b200.wrapper import MellowWrapper


class FakeEngine:
    """Mimics the Engine surface generate() uses; tokens are a pure function of (crop start of both clips, first prompt id)."""
    def __init__(self, max_batch, max_new_tokens):
        self.max_batch, self.max_new_tokens, self.device = max_batch, max_new_tokens, torch.device("cpu")
        self.options = {}

    def set_option(self, name, value):
        self.options[name] = value

    def close(self):
        pass

    def prepare_clip(self, pcm, sr, target_sr, resample, rng, out=None):
        ch, n_in = pcm.shape
        total = ch * (resampled_length(n_in, sr, target_sr) if resample and sr != target_sr else n_in)
        out[:] = float(plan_fit(total, S.CLIP_SAMPLES, rng))          # the draw (0 when the clip is tiled)
        return out

    def generate(self, a1, a2, ids, max_len, temperature=1.0, top_p=0.8, eos_id=0):
        b = a1.shape[0]
        key = (a1[:, 0].long() + 3 * a2[:, 0].long() + ids[:, 0].long()) % 997
        t = torch.arange(max_len)[None, :]
        toks = ((key[:, None] * 7 + t * 13) % 200 + 300).to(torch.int32)  # 300..499: printable stand-in bytes, never a stop id
        stop_at = (key % 5 + 2)[:, None]                                  # every row emits the stop id once, at step 2..6
        if eos_id >= 0:
            toks = torch.where(t == stop_at, torch.tensor(eos_id, dtype=torch.int32), toks)
            steps = int(stop_at.max()) + 1                                # the pass stops when all ITS rows have stopped
            return toks[:, :min(steps, max_len)]
        return torch.where(t == stop_at, torch.tensor(17, dtype=torch.int32), toks)     # split run: 17 = the custom stop id


def make_wrapper(cap):
    mw = object.__new__(MellowWrapper)
    mw.args = type("A", (), {"data": {"sampling_rate": 32000, "text_tokenization_len": 129}})()
    mw.tokenizer = ByteStandInTokenizer()
    mw.shard, mw.policy, mw.device = None, "split24", 0
    mw._fixed_batch, mw._fixed_new = cap, None
    mw.model = FakeEngine(cap, 300)
    mw._new_engine = lambda b, n: FakeEngine(b, n)
    return mw


def write_wavs(tmp, n):
    """Clips of different lengths and rates: some are tiled (no draw), some cropped (one draw each)."""
    paths = []
    for i in range(n):
        sr = [32000, 44100, 48000][i % 3]
        seconds = [4.0, 10.5, 12.0, 9.0, 11.25][i % 5]
        p = os.path.join(tmp, f"c{i}.wav")
        with wavmod.open(p, "wb") as f:
            f.setnchannels(1 + (i % 2)); f.setsampwidth(2); f.setframerate(sr)
            f.writeframes(np.zeros(int(sr * seconds) * (1 + (i % 2)), dtype="<i2").tobytes())
        paths.append(p)
    return paths


def examples_for(tmp, n):
    paths = write_wavs(tmp, 2 * n)
    return [[paths[i], paths[n + i], f"question {i}?"] for i in range(n)]


def run_generate(tmp, n, cap, stop_token):
    random.seed(123)
    return make_wrapper(cap).generate(examples_for(tmp, n), max_len=12, top_p=0.8, temperature=1.0, stop_token=stop_token)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, n, cap, stop_token, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    out = run_generate(tmp, n, cap, stop_token)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _sharded(tmp, n, cap, stop_token):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tmp, n, cap, stop_token, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return results


def test_sharded_generate_equals_single_process(tmp_path):
    tmp = str(tmp_path)
    n = 7
    single = run_generate(tmp, n, 128, "<|endoftext|>")
    assert len(single) == n and len(set(single)) > 1
    got = _sharded(tmp, n, 128, "<|endoftext|>")
    assert got[0] == single and got[1] == single                      # every rank returns the full list, in example order


def test_passes_and_custom_stop_token_keep_the_global_stop_rule(tmp_path):
    """Capacity 3 -> three passes for 7 examples; with a stop token other than '<|endoftext|>' the reference returns every
    row up to the GLOBAL stop step (wrapper.py:247-254), so passes / ranks must not stop on their own."""
    tmp = str(tmp_path)
    one_pass = run_generate(tmp, 7, 128, "!")                          # '!' = id 17 in the stand-in vocabulary
    three_passes = run_generate(tmp, 7, 3, "!")
    assert one_pass == three_passes
    got = _sharded(tmp, 7, 3, "!")
    assert got[0] == one_pass and got[1] == one_pass
    assert run_generate(tmp, 7, 3, "<|endoftext|>") == run_generate(tmp, 7, 128, "<|endoftext|>")
