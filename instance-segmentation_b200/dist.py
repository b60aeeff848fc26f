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


def decode_output_sharded(inputs, outs, infos, transforms, decode_cfg, device, decode_fn=None):
    """decode_output (utils/decode.py:444-461) of a GLOBAL batch over all ranks: every rank is handed the same global
    arguments (host tensors, e.g. pinned), decodes its contiguous shard of the images (shard_range; images are
    independent, utils/parell_util.py:5-8 - no collective on the data path) and the variable-length per-image results
    are gathered to rank 0 (gather_object).  Returns the full list (global image order) on rank 0, None elsewhere.
    `decode_fn` defaults to the drop-in utils.decode.decode_output."""
    if decode_fn is None:
        from .utils.decode import decode_output as decode_fn
    kp_out, regression, classification, anchors = outs
    n = kp_out[0].shape[0]
    rank, world = (dist.get_rank(), dist.get_world_size()) if is_dist() else (0, 1)
    lo, hi = shard_range(n, rank, world)
    mine = []
    if hi > lo:
        sub = (tuple(None if t is None else t[lo:hi] for t in kp_out), regression[lo:hi], classification[lo:hi], anchors)
        sub_inputs = inputs[lo:hi] if inputs.shape[0] == n else inputs
        mine = decode_fn(sub_inputs, sub, infos[lo:hi], transforms, decode_cfg, device)
    if world == 1:
        return mine
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, dst=0)
    if rank != 0:
        return None
    return [d for part in parts for d in part]
