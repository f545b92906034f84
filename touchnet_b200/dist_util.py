"""Small torch.distributed helpers for the data-parallel (FSDP2, weak-scaling) form of the path.

The hot path shards by packed rows: every rank featurises and trains on its own rows, there is no data-path
collective of our own - parameter all-gather / gradient reduce-scatter are FSDP2's (NCCL), exactly as the reference
applies them (ref: touchnet/models/helper_func.py:134-202).  What is ours on the host side is only the bookkeeping
below: per-rank batch seeds (ref: touchnet/data/datapipe.py:64-68 shards samples by dp_rank) and the max-over-ranks
timing / whole-job token accounting used by bench.py."""
from __future__ import annotations

import torch


def rank_seed(base_seed: int, rank: int) -> int:
    """Rank r draws its packed rows from seed base + r."""
    return int(base_seed) + int(rank)


def max_over_ranks(value: float, device, group=None) -> float:
    """Device time of a multi-GPU step = the slowest rank's."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(value: float, device, group=None) -> float:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def whole_job_tokens_per_s(tokens_per_rank_step: int, steps: int, world: int, ms_max: float) -> float:
    """tokens of ALL ranks / max-over-ranks time (pads included, the reference's own count: train.py:345)."""
    return tokens_per_rank_step * world * steps / (ms_max * 1e-3)
