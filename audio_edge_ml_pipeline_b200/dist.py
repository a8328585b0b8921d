"""Multi-GPU plumbing: clips are independent, so ranks shard them in contiguous blocks and no
collective touches the data path.  ``torch.distributed`` (NCCL on GPUs, gloo in the
CPU tests) is used only for the start barrier, the max-over-ranks timing and an optional
checksum gather."""

from __future__ import annotations

import os
from typing import Tuple


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Rank ``rank`` owns items ``[lo, hi)``; blocks are contiguous so that concatenating the
    ranks' outputs preserves loader order (base.py:199,219)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return n_items * rank // world, n_items * (rank + 1) // world


def env_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: str, device=None):
    import torch.distributed as dist
    world, rank, _ = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return world, rank


def barrier() -> None:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device="cpu") -> float:
    """Timing rule: a multi-GPU number is the MAX over ranks."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_per_rank: int, seconds_this_rank: float, device="cpu") -> float:
    """Whole-job throughput = units all ranks processed / slowest rank's time."""
    total = sum_over_ranks(float(units_per_rank), device)
    return total / max_over_ranks(seconds_this_rank, device)
