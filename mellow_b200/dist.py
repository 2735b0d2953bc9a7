"""Multi-GPU plumbing: one process per GPU, batch-sharded replicas.

Every row of the batch is independent on this path (eval-mode BatchNorm, per-token norms, per-row attention), so the
only collective is ONE broadcast of the packed weight arena at init: rank 0 loads / packs the checkpoint, every rank
receives the same bytes over NCCL (NVLink/NVSwitch) and binds them.  No per-step collective exists.
"""
import os

import torch
import torch.distributed as dist

from .engine import Engine


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_bounds(n_items, rank, world):
    """Contiguous slice [lo, hi) of ``n_items`` owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_arena(arena_cpu_or_none, nbytes, device, src=0):
    """rank ``src`` passes the packed CPU arena, the others None; returns the device arena on every rank."""
    if arena_cpu_or_none is not None:
        arena = arena_cpu_or_none.to(device)
    else:
        arena = torch.empty(nbytes, dtype=torch.uint8, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(arena, src=src)
    return arena


def gather_rows(local_rows, n_total, rank, world):
    """Host-side concatenation of per-rank int32 token matrices (all padded to the same width) in example order."""
    if world == 1:
        return local_rows
    lens = [shard_bounds(n_total, r, world) for r in range(world)]
    width = local_rows.shape[1]
    pad = max(hi - lo for lo, hi in lens)
    buf = torch.zeros(pad, width, dtype=local_rows.dtype, device=local_rows.device)
    buf[:local_rows.shape[0]] = local_rows
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, lens)], dim=0)


def build_engine(state_dict_fn, max_batch, max_new_tokens, policy="split"):
    """Create this rank's Engine; only rank 0 calls ``state_dict_fn()`` and packs, the arena is then broadcast."""
    rank, local_rank, world = init_process_group()
    eng = Engine(None, device=local_rank, max_batch=max_batch, max_new_tokens=max_new_tokens, policy=policy)
    arena_cpu = eng.pack_arena(state_dict_fn()) if rank == 0 else None
    eng.bind_arena(broadcast_arena(arena_cpu, eng.arena_bytes(), eng.device))
    return eng, rank, world
