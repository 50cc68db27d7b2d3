"""KL[q || p] for Gaussian q, p (mirrors gpflow/kullback_leiblers.py:31-165)."""
from __future__ import annotations

from typing import Any, Optional

from . import config, covariances, ops
from .inducing_variables import InducingVariables
from .kernels import Kernel
from .utilities.multipledispatch import Dispatcher

prior_kl = Dispatcher("prior_kl")


@prior_kl.register(InducingVariables, Kernel, object, object)
def _prior_kl(inducing_variable, kernel, q_mu, q_sqrt, whiten: bool = False):
    if whiten:
        return gauss_kl(q_mu, q_sqrt, None)                                   # :45-46
    K = covariances.Kuu(inducing_variable, kernel, jitter=config.default_jitter())  # :48
    return gauss_kl(q_mu, q_sqrt, K)


def gauss_kl(q_mu, q_sqrt, K=None, *, K_cholesky=None):
    """Sum over the L columns of KL[N(q_mu, q_sqrt q_sqrt^T) || N(0, K)] -> device fp64 scalar [1].
    q_mu [M, L]; q_sqrt [L, M, M] (lower-triangular part used) or [M, L] (diagonal); K / K_cholesky
    [M, M] or [L, M, M] or None (white)."""
    if (K is not None) and (K_cholesky is not None):
        raise ValueError("Ambiguous arguments: gauss_kl() must only be passed one of `K` or `K_cholesky`.")
    T = ops.torch()
    q_mu, q_sqrt = ops.to_device(q_mu), ops.to_device(q_sqrt)
    is_white = (K is None) and (K_cholesky is None)
    is_diag = q_sqrt.dim() == 2
    M, L = q_mu.shape
    acc = ops.zeros_scalar(1)  # accumulates twoKL
    # logdet of q: - sum log diag(Lq)^2   (:130)
    if is_diag:
        ops.reduce(ops.SUMLOGSQ, q_sqrt, M * L, 1, scale=-1.0, out=acc, accumulate=True)
    else:
        for l in range(L):
            ops.reduce(ops.SUMLOGSQ, q_sqrt[l], M, M + 1, scale=-1.0, out=acc, accumulate=True)
    if is_white:
        ops.reduce(ops.SUMSQ, q_mu, M * L, 1, out=acc, accumulate=True)                       # :124
        if is_diag:
            ops.reduce(ops.SUMSQ, q_sqrt, M * L, 1, out=acc, accumulate=True)                 # :134
        else:
            ops.tril_sumsq(q_sqrt, out=acc, accumulate=True)                                   # :120,134
        const = -float(M * L)                                                                  # :127
        return _finish(acc, const)
    # non-white
    if K is not None:
        K = ops.to_device(K)
        batched = K.dim() == 3
        chols = [ops.cholesky(K[l]) for l in range(L)] if batched else [ops.cholesky(K)]     # :107
    else:
        Kc = ops.to_device(K_cholesky)
        batched = Kc.dim() == 3
        chols = [(Kc[l], None) for l in range(L)] if batched else [(Kc, None)]
    for l in range(L):
        Lp, dinv = chols[l] if batched else chols[0]
        alpha = ops.trsm(Lp, ops.copy(q_mu[:, l:l + 1]), dinv=dinv)                           # :114
        ops.reduce(ops.SUMSQ, alpha, M, 1, out=acc, accumulate=True)                          # :124
        if is_diag and not batched:
            pass  # handled below once
        else:
            if is_diag:
                Lq = ops.full((M, M), 0.0, like=q_mu)
                ops.add_diag_(Lq, 0.0, ops.copy(q_sqrt[:, l:l + 1]).reshape(-1))
            else:
                Lq = ops.tril_(ops.copy(q_sqrt[l]))
            LpiLq = ops.trsm(Lp, Lq, dinv=dinv)                                               # :152
            ops.reduce(ops.SUMSQ, LpiLq, M * M, 1, out=acc, accumulate=True)                  # :153
        if batched:
            ops.reduce(ops.SUMLOGSQ, Lp, M, Lp.stride(0) + 1, out=acc, accumulate=True)       # :159-163
    if is_diag and not batched:                                                               # :136-145
        Lp, dinv = chols[0]
        eye = ops.add_diag_(ops.full((M, M), 0.0, like=q_mu), 1.0)
        Lp_inv = ops.trsm(Lp, eye, dinv=dinv)
        kinv = ops.colsumsq(Lp_inv)  # diag(Lp^-T Lp^-1)
        for l in range(L):
            w = ops.copy(q_sqrt[:, l:l + 1]).reshape(-1)
            tmp = ops.copy(w.reshape(M, 1))
            # sum_m kinv[m] * q[m]^2 = || sqrt-free form: scale rows then dot
            ops.scale_rows_(tmp, w)           # q^2
            ops.scale_rows_(tmp, kinv)        # kinv * q^2
            ops.reduce(ops.SUM, tmp, M, 1, out=acc, accumulate=True)
    if not batched:
        Lp, _ = chols[0]
        ops.reduce(ops.SUMLOGSQ, Lp, M, Lp.stride(0) + 1, scale=float(L), out=acc, accumulate=True)  # :162-163
    return _finish(acc, -float(M * L))


def _finish(acc, const: float):
    """0.5 * (acc + const) as a device scalar."""
    out = ops.full((1,), const, like=acc)
    ops.axpby(0.5, acc, 0.5, out)
    return out
