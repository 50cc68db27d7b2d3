"""Sharding of the hot path over the GPUs of one box.

The path partitions into independent units — minibatches of `SVGP.elbo`, latent GPs
(`SVGP.elbo(latent_range=...)`), independent outputs / replicas — and the only exchange is ONE
all-reduce (sum) of the fp64 objective scalar per evaluation (NCCL over NVLink on GPUs, gloo in the CPU
tests).  No data-path collective exists; a distributed Cholesky is out of scope."""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of `n_units` units: rank r gets [begin, end); the first
    `n_units % world` ranks get one extra unit.  Empty ranges are possible when world > n_units."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_sum_(scalar, group=None):
    """In-place sum over ranks of a 1-element fp64 tensor; identity when torch.distributed is not
    initialised (single GPU)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(scalar, op=dist.ReduceOp.SUM, group=group)
    return scalar


def sharded_sum(evaluate: Callable[[int, int], "object"], n_units: int, rank: Optional[int] = None,
                world: Optional[int] = None, group=None):
    """`evaluate(begin, end)` returns this rank's partial objective (1-element fp64 tensor, zeros for an
    empty range); the partial sums are all-reduced.  Used for latent-GP sharding of SVGP.elbo and for
    sums over independent outputs."""
    import torch.distributed as dist

    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    b, e = shard_range(n_units, rank, world)
    return allreduce_sum_(evaluate(b, e), group)
