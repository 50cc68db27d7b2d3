"""Linear kernel (mirrors gpflow/kernels/linears.py:25-68)."""
from __future__ import annotations

from typing import Any

from .. import _lib
from ..base import Parameter, positive
from .base import ActiveDims, Kernel


class Linear(Kernel):
    def __init__(self, variance: Any = 1.0, active_dims: ActiveDims = None) -> None:
        super().__init__(active_dims)
        self.variance = Parameter(variance, transform=positive())
        self._validate_ard_active_dims(self.variance)

    @property
    def ard(self) -> bool:
        return self.variance.numpy().ndim > 0

    def _leaf_record(self, D: int) -> dict:
        v = self.variance.numpy()
        if v.ndim > 0:
            return {"op": _lib.K_LINEAR, "ard": v.reshape(-1)}
        return {"op": _lib.K_LINEAR, "variance": float(v)}


class Polynomial(Linear):
    """(sigma^2 x.y + offset)^degree  (gpflow/kernels/linears.py:71-112)."""

    def __init__(self, degree: Any = 3.0, variance: Any = 1.0, offset: Any = 1.0, active_dims: ActiveDims = None) -> None:
        super().__init__(variance, active_dims)
        self.degree = degree
        self.offset = Parameter(offset, transform=positive())

    def _leaf_record(self, D: int) -> dict:
        rec = super()._leaf_record(D)
        rec["op"] = _lib.K_POLYNOMIAL
        rec["lengthscale"] = float(self.offset.numpy())   # gpk_knode: the offset rides in the `lengthscale` field
        rec["alpha"] = float(self.degree)                 # and the degree in `alpha`
        return rec
