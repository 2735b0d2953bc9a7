"""CPU: the N>1 plumbing (batch sharding, one weight broadcast at init, host gather of token ids) with gloo, world 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mellow_b200 import dist as mdist


def test_shard_bounds_cover_every_example_once():
    for n in (1, 2, 7, 128, 256, 513):
        for world in (1, 2, 4, 8):
            spans = [mdist.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    mdist.init_process_group(backend="gloo")
    nbytes = 4096
    src = (torch.arange(nbytes, dtype=torch.int64) % 251).to(torch.uint8) if rank == 0 else None
    arena = mdist.broadcast_arena(src, nbytes, torch.device("cpu"))
    ok_arena = bool(torch.equal(arena, (torch.arange(nbytes, dtype=torch.int64) % 251).to(torch.uint8)))
    lo, hi = mdist.shard_bounds(n_total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.int32)[:, None] * 10 + torch.arange(3, dtype=torch.int32)[None, :]
    full = mdist.gather_rows(local, n_total, rank, world)
    want = torch.arange(n_total, dtype=torch.int32)[:, None] * 10 + torch.arange(3, dtype=torch.int32)[None, :]
    q.put((rank, ok_arena, bool(torch.equal(full, want))))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(r[1] and r[2] for r in results)
