"""Static kernels (mirrors gpflow/kernels/statics.py:25-91)."""
from __future__ import annotations

from typing import Any

from .. import _lib
from ..base import Parameter, positive
from .base import ActiveDims, Kernel


class Static(Kernel):
    _op = -1

    def __init__(self, variance: Any = 1.0, active_dims: ActiveDims = None) -> None:
        super().__init__(active_dims)
        self.variance = Parameter(variance, transform=positive())

    def _leaf_record(self, D: int) -> dict:
        return {"op": self._op, "variance": float(self.variance.numpy())}


class White(Static):
    """diag(sigma^2) iff X2 is None, zeros otherwise (statics.py:57-63)."""

    _op = _lib.K_WHITE


class Constant(Static):
    _op = _lib.K_CONSTANT


Bias = Constant
