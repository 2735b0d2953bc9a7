"""GPU, two devices: ``MellowWrapper.generate()`` sharded over two ranks (NCCL, one process per GPU, one weight
broadcast is not even needed here: every rank packs the same synthetic checkpoint) must return, on every rank, the
list the single-GPU run returns.  Skipped on boxes with fewer than two GPUs (run with ``gpurun --gpus 2``)."""
import os
import random
import socket
import wave as wavmod

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_examples(tmp, n):
    """Real wav files of mixed rate / length (tile and crop branches) from the seeded synthetic waveforms."""
    from mellow_b200 import synth
    w = synth.synthetic_waveforms(2 * n, seed=555)
    paths = []
    for i in range(2 * n):
        sr = [32000, 44100][i % 2]
        seconds = [10.0, 11.0, 6.5][i % 3]
        x = torch.nn.functional.interpolate(w[i][None, None, :], size=int(sr * seconds), mode="linear")[0, 0]
        p = os.path.join(tmp, f"clip{i}.wav")
        with wavmod.open(p, "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(sr)
            f.writeframes((x.numpy() * 32767.0).round().astype("<i2").tobytes())
        paths.append(p)
    return [[paths[i], paths[n + i], f"what is different? ({i})"] for i in range(n)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, examples, max_len, q):
    import torch.distributed as dist
    from mellow_b200 import MellowWrapper
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group(backend="nccl", rank=rank, world_size=world)
    mw = MellowWrapper(config="v0", model="v0", device=rank, use_cuda=True, checkpoint="synthetic")
    random.seed(7)
    out = mw.generate(examples=examples, max_len=max_len, top_p=0.8, temperature=1.0)
    q.put((rank, out))
    dist.barrier()
    mw.model.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_generate_on_two_gpus_equals_one_gpu(tmp_path):
    import torch.multiprocessing as mp
    from mellow_b200 import MellowWrapper
    n, max_len = 5, 24
    examples = _write_examples(str(tmp_path), n)
    mw = MellowWrapper(config="v0", model="v0", device=0, use_cuda=True, checkpoint="synthetic")
    random.seed(7)
    single = mw.generate(examples=examples, max_len=max_len, top_p=0.8, temperature=1.0)
    mw.model.close()
    del mw
    assert len(single) == n
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, examples, max_len, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert results[0] == single and results[1] == single
