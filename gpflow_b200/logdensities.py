"""Log densities on the hot path (mirrors gpflow/logdensities.py:29-30,139-156)."""
from __future__ import annotations

import math

from . import ops


def multivariate_normal(x, mu, L, dinv=None):
    """log N(x | mu, L L^T) per column -> device fp64 tensor [P] (logdensities.py:139-156)."""
    x, L = ops.to_device(x), ops.to_device(L)
    d = ops.copy(x)
    if mu is not None:
        ops.axpby(-1.0, ops.to_device(mu), 1.0, d)
    alpha = ops.trsm(L, d, dinv=dinv)                      # :150
    n, P = alpha.shape
    T = ops.torch()
    out = T.empty((P,), dtype=T.float64, device=alpha.device)
    logdet = ops.reduce(ops.SUMLOG, L, n, L.stride(0) + 1)  # :154
    for p in range(P):
        ops.reduce(ops.SUMSQ, alpha[:, p:p + 1], n, alpha.stride(0), scale=-0.5, out=out[p:p + 1])  # :152
        ops.axpby(-1.0, logdet, 1.0, out[p:p + 1])
    const = ops.full((P,), -0.5 * n * math.log(2 * math.pi), like=out)  # :153
    ops.axpby(1.0, const, 1.0, out)
    return out
