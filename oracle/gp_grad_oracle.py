"""Gradient oracle for SURVEY.md 8(f) rank 1 (test infrastructure, like gp_oracle.py; not imported by the product).

The reference obtains d(LML)/d(theta) from TensorFlow autodiff through gpflow/models/gpr.py:91-107
(`training_loss_closure`, gpflow/models/training_mixins.py:43-78).  The closed form restated here is the standard

    d LML / d theta = 1/2 sum_p alpha_p^T (dK/dtheta) alpha_p - P/2 tr(K^-1 dK/dtheta),   alpha = K^-1 (Y - m),

with K = kernel(X) + sigma_n^2 I, for an isotropic stationary kernel (scalar variance and lengthscale) and the
likelihood variance.  It is pinned by central finite differences of gp_oracle.gpr_log_marginal_likelihood
(tests/test_oracle_grad.py): this is the target the device-side backward pass of the next round has to match.
"""
from __future__ import annotations

from typing import Dict, Tuple, Union

import numpy as np

from . import gp_oracle as O


def stationary_dK(kernel: O.Stationary, X: np.ndarray) -> Dict[str, np.ndarray]:
    """dK/d(variance) and dK/d(lengthscales) of a stationary kernel on X; with an ARD lengthscale vector the entry
    "lengthscales" is a [D_active, N, N] stack.  With s = sum_d ((x_d - x'_d) / l_d)^2 and r = sqrt(s):
    dk/dl_d = (dk/ds) (-2 (x_d - x'_d)^2 / l_d^3).  (stationaries.py:209-210, 250-251, 270-271, 290-292, 311-313)"""
    var = float(np.asarray(kernel.variance))
    ell = np.asarray(kernel.lengthscales, dtype=np.float64)
    K = kernel(X)
    Xs = kernel.slice(X)[0]
    r2 = np.maximum(kernel.scaled_squared_euclid_dist(Xs), 0.0)
    r = np.sqrt(r2)
    with np.errstate(divide="ignore", invalid="ignore"):
        if isinstance(kernel, O.SquaredExponential):
            dkds = -0.5 * K
        elif isinstance(kernel, O.Exponential):
            dkds = np.where(r > 0, -K / (4.0 * r), 0.0)
        elif isinstance(kernel, O.Matern12):
            dkds = np.where(r > 0, -K / (2.0 * r), 0.0)
        elif isinstance(kernel, O.Matern32):
            dkds = -1.5 * var * np.exp(-np.sqrt(3.0) * r)
        elif isinstance(kernel, O.Matern52):
            s5 = np.sqrt(5.0)
            dkds = -(5.0 / 6.0) * var * (1.0 + s5 * r) * np.exp(-s5 * r)
        else:
            raise NotImplementedError(type(kernel).__name__)
    if ell.ndim == 0:
        dl = dkds * (-2.0 * r2 / float(ell))
    else:
        diff2 = (Xs[:, None, :] - Xs[None, :, :]) ** 2            # [N, N, D]
        dl = np.stack([dkds * (-2.0 * diff2[:, :, d] / ell[d] ** 3) for d in range(ell.shape[0])])
    return {"variance": K / var, "lengthscales": dl}


def gpr_lml_and_grad(X: np.ndarray, Y: np.ndarray, kernel: O.Stationary, noise_variance: float,
                     mean_function=None) -> Tuple[float, Dict[str, float]]:
    """LML (gpr.py:91-107) and its gradient w.r.t. kernel variance, lengthscale and likelihood variance."""
    N, P = Y.shape
    K = O.add_noise_cov(kernel(X), noise_variance)
    L = O.cholesky(K)
    err = Y - O._mean(mean_function, X, P)
    lml = float(np.sum(O.multivariate_normal(Y, O._mean(mean_function, X, P), L)))
    alpha = O.tri_solve(L, O.tri_solve(L, err), trans=True)             # K^-1 err
    Linv = O.tri_solve(L, np.eye(N, dtype=X.dtype))
    Kinv = Linv.T @ Linv
    G = 0.5 * (alpha @ alpha.T - P * Kinv)                               # dLML/dK
    dK = stationary_dK(kernel, X)
    grad = {"variance": float(np.sum(G * dK["variance"]))}
    dl = dK["lengthscales"]
    grad["lengthscales"] = float(np.sum(G * dl)) if dl.ndim == 2 else np.array([np.sum(G * d) for d in dl])
    grad["noise_variance"] = float(np.trace(G))
    return lml, grad
