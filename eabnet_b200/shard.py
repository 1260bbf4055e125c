"""Host-side plumbing of the multi-GPU inference path: utterance sharding and max-over-ranks timing.

The path has no exchange step: utterances are independent end to end, so ranks only
agree on who enhances what and on the slowest rank's time."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of `n_items` utterances for `rank` (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def max_over_ranks(values, device=None):
    """All-reduce MAX of a list of floats (returns the list unchanged when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return list(values)
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]
