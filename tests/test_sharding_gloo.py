"""world_size-2 gloo tests (CPU) of the N>1 path's host logic: unit partitioning and the single scalar
all-reduce.  Partial objectives come from the oracle (test infrastructure) since there is no GPU here."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpflow_b200.sharding import shard_range, sharded_sum


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                b, e = shard_range(n, r, world)
                assert 0 <= b <= e <= n and e - b in (n // world, n // world + 1)
                cover += list(range(b, e))
            assert cover == list(range(n))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gp_oracle as O

    rng = np.random.default_rng(0)
    N, M, P, D = 120, 17, 5, 3
    X, Y, Z = rng.standard_normal((N, D)), rng.standard_normal((N, P)), rng.standard_normal((M, D))
    k = O.SquaredExponential(lengthscales=1.3) + O.White(variance=0.1)
    q_mu = rng.standard_normal((M, P))
    q_sqrt = np.stack([np.tril(rng.standard_normal((M, M))) * 0.2 + np.eye(M) for _ in range(P)])

    def latent_share(b, e):  # the data term and KL of latents [b, e) — what gpk_svgp_elbo(p_begin, p_end) returns
        if e <= b:
            return torch.zeros(1, dtype=torch.float64)
        v = O.svgp_elbo(X, Y[:, b:e], Z, k, q_mu[:, b:e], q_sqrt[b:e], 0.2, whiten=True, num_data=1000)
        return torch.tensor([v], dtype=torch.float64)

    total = sharded_sum(latent_share, P)                       # latent sharding: shares sum to the ELBO
    full = O.svgp_elbo(X, Y, Z, k, q_mu, q_sqrt, 0.2, whiten=True, num_data=1000)

    def minibatch(b, e):                                       # independent minibatches: one per unit
        s = sum(O.svgp_elbo(X[i * 30:(i + 1) * 30], Y[i * 30:(i + 1) * 30], Z, k, q_mu, q_sqrt, 0.2, num_data=1000)
                for i in range(b, e))
        return torch.tensor([s], dtype=torch.float64)

    mb_total = sharded_sum(minibatch, 4)
    mb_ref = sum(O.svgp_elbo(X[i * 30:(i + 1) * 30], Y[i * 30:(i + 1) * 30], Z, k, q_mu, q_sqrt, 0.2, num_data=1000)
                 for i in range(4))
    q.put((rank, float(total), full, float(mb_total), mb_ref))
    dist.barrier()
    dist.destroy_process_group()


def test_latent_and_minibatch_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, total, full, mb_total, mb_ref in res:
        np.testing.assert_allclose(total, full, rtol=1e-12)     # every rank holds the all-reduced ELBO
        np.testing.assert_allclose(mb_total, mb_ref, rtol=1e-12)


class _OracleSVGP:
    """Stands in for gpflow_b200.models.SVGP on CPU: same sharding-facing interface (`elbo` with `latent_range`,
    `batch_total`, `include_kl`), values from the oracle."""

    def __init__(self, O, Z, k, q_mu, q_sqrt, noise, num_data):
        self.O, self.Z, self.k, self.q_mu, self.q_sqrt, self.noise, self.num_data = O, Z, k, q_mu, q_sqrt, noise, num_data
        self.num_latent_gps = q_mu.shape[1]
        self.whiten = True

    def elbo(self, data, *, latent_range=None, batch_total=None, include_kl=True):
        O = self.O
        X, Y = (np.asarray(d) for d in data)
        p0, p1 = (0, self.num_latent_gps) if latent_range is None else latent_range
        B = X.shape[0] if batch_total is None else batch_total
        # ELBO = sum_rows var_exp * num_data / B - KL  (gpflow/models/svgp.py:173-181): evaluate with num_data chosen so
        # that the oracle's own scale num_data / X.shape[0] equals num_data / B
        nd = self.num_data * X.shape[0] / B
        v = O.svgp_elbo(X, Y[:, p0:p1], self.Z, self.k, self.q_mu[:, p0:p1], self.q_sqrt[p0:p1], self.noise, whiten=True,
                        num_data=nd)
        if not include_kl:
            v += O.gauss_kl(self.q_mu[:, p0:p1], self.q_sqrt[p0:p1], None)
        return torch.tensor(v, dtype=torch.float64)


def _worker_modes(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpflow_b200 import sharding
    from oracle import gp_oracle as O

    rng = np.random.default_rng(1)
    N, M, P, D = 90, 13, 5, 3
    X, Y, Z = rng.standard_normal((N, D)), rng.standard_normal((N, P)), rng.standard_normal((M, D))
    k = O.SquaredExponential(lengthscales=1.1) + O.White(variance=0.1)
    q_mu = rng.standard_normal((M, P))
    q_sqrt = np.stack([np.tril(rng.standard_normal((M, M))) * 0.2 + np.eye(M) for _ in range(P)])
    model = _OracleSVGP(O, Z, k, q_mu, q_sqrt, 0.2, 5000)
    full = O.svgp_elbo(X, Y, Z, k, q_mu, q_sqrt, 0.2, whiten=True, num_data=5000)
    rows = float(sharding.svgp_elbo_row_sharded(model, (X, Y)))
    lat = float(sharding.svgp_elbo_latent_sharded(model, (X, Y), shard_solve=False))
    q.put((rank, rows, lat, full))
    dist.barrier()
    dist.destroy_process_group()


def test_row_and_latent_sharded_elbo_world2():
    """The shares of the row-sharded and latent-sharded evaluations of ONE minibatch sum to the full ELBO
    (gpflow/models/svgp.py:173-181; rank 0 alone carries the KL in the row-sharded mode)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_modes, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rows, lat, full in res:
        np.testing.assert_allclose(rows, full, rtol=1e-11)
        np.testing.assert_allclose(lat, full, rtol=1e-11)
