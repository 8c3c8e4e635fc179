"""Data-parallel plumbing for the KD step: one process per GPU, utterances sharded over ranks, student gradients
summed by ONE NCCL all-reduce over NVLink (optim.FlatBuffers.all_reduce) -- nothing else is multi-GPU on this path.

Mirrors what accelerate/DDP do implicitly for ref:training/run_distillation.py:1607-1614.  Loss normalisation stays
per-rank (mean over the local unmasked tokens) followed by the gradient average, as under DDP.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns
    (rank, local_rank, world_size); a plain `python` launch gives (0, 0, 1) without creating a process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [begin, end) slice of `n_items` utterances owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Slice every tensor of a global batch along dim 0 for this rank."""
    n = next(iter(batch.values())).shape[0]
    b, e = shard_range(n, rank, world)
    return {k: v[b:e] for k, v in batch.items()}


def broadcast_parameters(model, src: int = 0):
    """All ranks start from rank `src`'s weights (what DDP does at construction)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src=src)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host float over ranks (used for the bench's max-over-ranks step time)."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
