"""Mean functions Zero / Constant / Linear (mirrors gpflow/functions.py:96-126,173-204)."""
from __future__ import annotations

from typing import Any

import numpy as np

from . import config, ops
from .base import Module, Parameter


class MeanFunction(Module):
    def __call__(self, X):
        raise NotImplementedError


class Zero(MeanFunction):
    def __init__(self, output_dim: int = 1) -> None:
        self.output_dim = output_dim

    def __call__(self, X):  # functions.py:201-204
        X = ops.to_device(X)
        return ops.full((X.shape[0], self.output_dim), 0.0, like=X)


class Constant(MeanFunction):
    def __init__(self, c: Any = None) -> None:
        c = np.zeros(1) if c is None else c
        self.c = Parameter(np.atleast_1d(np.asarray(c, dtype=config.default_float())))

    def __call__(self, X):  # functions.py:187-192
        X = ops.to_device(X)
        c = self.c.numpy()
        out = ops.empty((X.shape[0], c.shape[0]), like=X)
        for q in range(c.shape[0]):
            ops.fill(out[:, q:q + 1], float(c[q]))
        return out


class Linear(MeanFunction):
    def __init__(self, A: Any = None, b: Any = None) -> None:
        A = np.ones((1, 1), dtype=config.default_float()) if A is None else A
        b = np.zeros(1, dtype=config.default_float()) if b is None else b
        self.A = Parameter(np.atleast_2d(A))
        self.b = Parameter(np.atleast_1d(b))

    def __call__(self, X):  # functions.py:124-126
        X = ops.to_device(X)
        A = ops.to_device(self.A)
        out = ops.gemm(X, A)
        b = self.b.numpy()
        Q = out.shape[1]
        for q in range(Q):
            bq = float(b[q] if b.shape[0] > 1 else b[0])
            if bq != 0.0:
                col = out[:, q:q + 1]
                ones = ops.full(col.shape, bq, like=out)
                ops.axpby(1.0, ones, 1.0, col)
        return out
