"""Batch sharding across GPUs: image k of B goes to rank floor(k * G / B) -- contiguous blocks,
no collective on the data path.  torch.distributed is used only for the barrier and to reduce device times."""
from __future__ import annotations


def shard_range(n_images: int, rank: int, world: int) -> tuple[int, int]:
    """[first, last) image indices owned by `rank`: the k with floor(k * world / n_images) == rank."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    first = -(-rank * n_images // world)          # ceil(rank * n / world)
    last = -(-(rank + 1) * n_images // world)
    return first, last


def owner_of(k: int, n_images: int, world: int) -> int:
    return k * world // n_images


def job_time_ms(local_ms: float, device=None) -> float:
    """Whole-job device time = max over ranks (the slowest GPU defines the batch)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(local_ms)
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
