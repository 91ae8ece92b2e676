"""Multi-GPU = independent replicas: one process, one sequence and one handle per GPU, no data-path
collective.  torch.distributed is used only for the barrier and the max-over-ranks timing that bench.py reports.
These helpers are backend-agnostic so that the rank logic is covered by gloo tests on CPU."""
from __future__ import annotations

import os


def rank_info():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def sequence_seed_index(rank: int) -> int:
    """Rank r streams synthetic sequence r (seed 20241108 + r): BASELINE config 'independent sequences, one per GPU'."""
    return rank


def max_over_ranks(values, dist=None, device="cpu"):
    """Element-wise MAX of a list of floats over all ranks (device timings: the slowest rank defines the step time)."""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def aggregate_throughput(world: int, units_per_rank: int, max_ms: float) -> float:
    """Whole-job units/s: every rank processed units_per_rank units within the max-over-ranks time (weak scaling)."""
    return world * units_per_rank / (max_ms * 1e-3)
