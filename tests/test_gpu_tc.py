"""GPU tests of the tcgen05 (int8-sliced fp64) symmetric rank-k update against NumPy fp64, through
gpk_potrf's recursion (engine on) and directly via a Cholesky whose trailing updates use it."""
import os

import numpy as np
import pytest
import scipy.linalg as sla
from numpy.testing import assert_allclose

import gpflow_b200 as gpf
from gpflow_b200 import ops
from tests.helpers import to_np

pytestmark = pytest.mark.gpu


def _spd(n, rng, cond_shift=0.5):
    A = rng.standard_normal((n, n + 5))
    return A @ A.T / n + cond_shift * np.eye(n)


@pytest.mark.parametrize("n", [512, 640, 1000, 1536, 2048])
def test_potrf_with_tcgen05_trailing_update(cuda_device, n):
    """n >= 512 engages the tcgen05 path for the top level(s) of the recursion (K >= 256)."""
    rng = np.random.default_rng(n)
    K = _spd(n, rng)
    L, _ = ops.cholesky(ops.to_device(K))
    ref = sla.cholesky(K, lower=True)
    # digit truncation: K * S * 2^-49 relative to row maxima (S = 7) -> ~1e-11 on L
    assert_allclose(to_np(L), ref, rtol=2e-9, atol=2e-10)
    Ld = to_np(L)
    resid = np.linalg.norm(Ld @ Ld.T - K) / np.linalg.norm(K)
    assert resid < 1e-11, resid


def test_potrf_tc_extra_rows_and_scaling(cuda_device):
    """Rows with very different magnitudes exercise the per-row power-of-two scales."""
    rng = np.random.default_rng(7)
    n, p = 1024, 2
    D = np.exp(rng.uniform(-6, 6, n))                      # row/column scaling over 5 decades
    K = _spd(n, rng) * D[:, None] * D[None, :]
    Y = rng.standard_normal((n, p)) * D[:, None]
    A = np.zeros((n + p, n))
    A[:n] = np.tril(K)
    A[n:] = Y.T
    Ad = ops.to_device(A)
    ops.potrf(Ad, n)
    L = sla.cholesky(K, lower=True)
    got = to_np(Ad)
    assert_allclose(np.tril(got[:n]) / D[:, None], L / D[:, None], rtol=1e-7, atol=1e-9)
    alpha = sla.solve_triangular(L, Y, lower=True).T
    assert_allclose(got[n:], alpha, rtol=1e-7, atol=1e-8)


def test_gpr_lml_tc_vs_dmma_engines(cuda_device):
    """The fused GPR LML agrees between the tcgen05 engine and the DMMA engine to ~1e-10."""
    from oracle import gp_oracle as O

    d = O.make_data(2, 2048, 8, 1)
    k = gpf.kernels.Matern52(lengthscales=np.sqrt(8.0))
    m = gpf.models.GPR((d["X"], d["Y"]), k, noise_variance=0.1)
    lml = float(m.log_marginal_likelihood())
    ref = O.gpr_log_marginal_likelihood(d["X"], d["Y"], O.Matern52(lengthscales=np.sqrt(8.0)), 0.1)
    assert_allclose(lml, ref, rtol=1e-9)
