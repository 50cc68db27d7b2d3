"""Sharding of the hot path over the GPUs of one box.

The path partitions into independent units — minibatches of `SVGP.elbo`, rows of one minibatch, latent GPs
(`SVGP.elbo(latent_range=...)`), independent outputs / replicas — and the exchange is ONE all-reduce (sum) of the
fp64 objective scalar per evaluation (NCCL over NVLink on GPUs, gloo in the CPU tests).  Latent sharding
additionally shards the triangular solve A = Lm^-1 Kuf by minibatch columns and all-gathers A [M, B] (the only
data-path collective; without it every rank would repeat the whole solve).  A distributed Cholesky is out of scope."""
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


def svgp_elbo_row_sharded(model, data, rank: Optional[int] = None, world: Optional[int] = None, group=None):
    """ELBO of ONE minibatch with its ROWS partitioned over the ranks (gpflow/models/svgp.py:173-181: the data term is
    a sum over rows; conditionals/util.py:125-164: column n of Kuf / A / LTA depends on x_n only).  Every rank
    factorises Kuu; rank 0 adds the KL.  One scalar all-reduce."""
    import torch.distributed as dist

    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    X, Y = data
    B = int(X.shape[0])
    b, e = shard_range(B, rank, world)
    share = model.elbo((X[b:e], Y[b:e]), batch_total=B, include_kl=(rank == 0))
    return allreduce_sum_(share.reshape(1), group)


def svgp_elbo_latent_sharded(model, data, rank: Optional[int] = None, world: Optional[int] = None, group=None,
                             shard_solve: bool = True):
    """ELBO of ONE minibatch with the LATENT GPs partitioned over the ranks (every latent has its own q_mu[:, p],
    q_sqrt[p] and KL term; kullback_leiblers.py:72-74,124-134).  With `shard_solve` the triangular solve
    A = Lm^-1 Kuf, which all latents share, is sharded by minibatch columns and all-gathered (pack -> NCCL all-gather
    -> unpack); otherwise every rank repeats it.  One scalar all-reduce at the end."""
    import torch.distributed as dist
    from . import ops

    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    P = model.num_latent_gps
    p0, p1 = shard_range(P, rank, world)
    X, Y = data
    B = int(X.shape[0])
    if world == 1 or not shard_solve or not model.whiten or B % world != 0:
        if p1 <= p0:
            return allreduce_sum_(ops.zeros_scalar(1), group)
        return allreduce_sum_(model.elbo(data, latent_range=(p0, p1)).reshape(1), group)
    Bc = B // world
    A = model.solve_columns(data, (rank * Bc, (rank + 1) * Bc))      # [M, B] view, own columns valid
    M = int(A.shape[0])
    mine = ops.empty((M, Bc), like=A)
    ops.axpby(1.0, A[:, rank * Bc:(rank + 1) * Bc], 0.0, mine)          # pack (strided -> contiguous)
    gathered = ops.empty((world, M, Bc), like=A)
    dist.all_gather_into_tensor(gathered.view(-1), mine.view(-1), group=group)
    for r in range(world):
        if r != rank:
            ops.axpby(1.0, gathered[r], 0.0, A[:, r * Bc:(r + 1) * Bc])  # unpack
    if p1 <= p0:
        return allreduce_sum_(ops.zeros_scalar(1), group)
    return allreduce_sum_(model.elbo_from_columns(data, (p0, p1)).reshape(1), group)
