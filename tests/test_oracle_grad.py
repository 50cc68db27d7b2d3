"""Pins the gradient oracle (SURVEY.md 8(f) rank 1 target) by central finite differences of the LML oracle."""
import numpy as np
import pytest

from oracle import gp_grad_oracle as G
from oracle import gp_oracle as O


@pytest.mark.parametrize("cls", [O.SquaredExponential, O.Matern32, O.Matern52])
@pytest.mark.parametrize("P", [1, 2])
def test_gpr_lml_gradient_matches_finite_differences(cls, P):
    rng = np.random.default_rng(20220523)
    N, D = 60, 3
    X = rng.standard_normal((N, D))
    Y = np.sin(X[:, :1]) + 0.1 * rng.standard_normal((N, P))
    var, ell, s2 = 1.3, 1.7, 0.15
    lml, g = G.gpr_lml_and_grad(X, Y, cls(variance=var, lengthscales=ell), s2)
    assert abs(lml - O.gpr_log_marginal_likelihood(X, Y, cls(variance=var, lengthscales=ell), s2)) < 1e-10

    def f(v, l, s):
        return O.gpr_log_marginal_likelihood(X, Y, cls(variance=v, lengthscales=l), s)

    h = 1e-5
    fd = {"variance": (f(var + h, ell, s2) - f(var - h, ell, s2)) / (2 * h),
          "lengthscales": (f(var, ell + h, s2) - f(var, ell - h, s2)) / (2 * h),
          "noise_variance": (f(var, ell, s2 + h) - f(var, ell, s2 - h)) / (2 * h)}
    for k in fd:
        assert abs(g[k] - fd[k]) <= 2e-6 * max(1.0, abs(fd[k])), (k, g[k], fd[k])


@pytest.mark.parametrize("cls", [O.SquaredExponential, O.Matern12, O.Matern32, O.Matern52])
def test_stationary_kernel_derivatives_offdiagonal(cls):
    """dK/dl and dK/dvariance against finite differences of the kernel itself.  Off-diagonal entries only: on the
    diagonal the reference's norm-expansion distance (utilities/ops.py:109-111) leaves rounding noise in r^2 that
    jumps with l, so for the sqrt-like Matern12 a finite difference of K_ii is noise (in the reference too)."""
    rng = np.random.default_rng(3)
    X = rng.standard_normal((12, 3))
    var, ell, h = 1.3, 1.7, 1e-6
    dK = G.stationary_dK(cls(variance=var, lengthscales=ell), X)
    off = ~np.eye(12, dtype=bool)
    fd_l = (cls(variance=var, lengthscales=ell + h)(X) - cls(variance=var, lengthscales=ell - h)(X)) / (2 * h)
    fd_v = (cls(variance=var + h, lengthscales=ell)(X) - cls(variance=var - h, lengthscales=ell)(X)) / (2 * h)
    np.testing.assert_allclose(dK["lengthscales"][off], fd_l[off], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(dK["variance"][off], fd_v[off], rtol=1e-6, atol=1e-8)
