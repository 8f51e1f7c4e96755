"""Host-side plumbing for the multi-GPU path (one process per GPU).

The collectives themselves run inside libhammock_b200.so over NCCL; torch.distributed is used only
to hand rank 0's NCCL unique id to the other ranks (works with the gloo backend on CPU too, which
is how the rendezvous logic is tested without GPUs).
"""
from __future__ import annotations

from typing import Callable


def exchange_unique_id(dist, rank: int, make_id: Callable[[], bytes]) -> bytes:
    """Rank 0 creates the 128-byte id with make_id(); every rank returns the same bytes."""
    import torch
    backend = dist.get_backend()
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    if rank == 0:
        uid = make_id()
        if len(uid) != 128:
            raise ValueError("NCCL unique id must be 128 bytes")
        t = torch.tensor(list(uid), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(128, dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def shard_range(n: int, world: int, rank: int):
    """Contiguous share [lo, hi) of n work items (the split the library uses for the database stripes
    of phase 1 and the query shards of phase 2)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def merge_best_hits(per_rank_keys, per_rank_overflow, kb: int):
    """Host-side statement of what hmk_topk_merge does with the all-gathered per-rank lists
    (used by the CPU tests of the N > 1 path): the kb largest keys of the union, descending, and
    whether anything was dropped anywhere."""
    import numpy as np
    allk = np.concatenate([np.asarray(k, dtype=np.uint64) for k in per_rank_keys]) if per_rank_keys else np.zeros(0, np.uint64)
    allk = np.sort(allk)[::-1]
    overflow = bool(any(per_rank_overflow)) or len(allk) > kb
    return allk[:kb], overflow


def make_key(score: int, tierank: int) -> int:
    """hmk_key_make (csrc/hmk_common.h): bigger key == preferred partner."""
    return (((score & 0xFFFFFFFF) ^ 0x80000000) << 32) | ((~tierank) & 0xFFFFFFFF)
