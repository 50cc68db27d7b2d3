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
