"""Multi-GPU plumbing.  Images are independent (the reference maps decode_single over the batch,
utils/parell_util.py:5-8), so a batch shards into contiguous per-rank chunks with NO collective on the data
path; torch.distributed (NCCL on GPUs, gloo in the CPU tests) only reduces timings and gathers result counts."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous chunk [lo, hi) of `n_items` for `rank`; sizes differ by at most one, earlier ranks get the extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def barrier() -> None:
    if is_dist():
        dist.barrier()


def reduce_max(value: float, device=None) -> float:
    """max over ranks of a host scalar (step time: the job is as slow as its slowest rank)"""
    if not is_dist():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, device=None) -> float:
    if not is_dist():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_ints(values, device=None):
    """all_gather of a fixed-length int list; returns a list (per rank) of lists"""
    if not is_dist():
        return [list(values)]
    t = torch.tensor(list(values), dtype=torch.int64, device=device or "cpu")
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]
