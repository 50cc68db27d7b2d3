"""Stationary kernels (mirrors gpflow/kernels/stationaries.py:35-313)."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from .. import _lib
from ..base import Parameter, positive
from .base import ActiveDims, Kernel


class Stationary(Kernel):
    _op = -1

    def __init__(self, variance: Any = 1.0, lengthscales: Any = 1.0, **kwargs: Any) -> None:
        for kwarg in kwargs:  # stationaries.py:56-58
            if kwarg not in {"name", "active_dims"}:
                raise TypeError(f"Unknown keyword argument: {kwarg}")
        super().__init__(**kwargs)
        self.variance = Parameter(variance, transform=positive())
        self.lengthscales = Parameter(lengthscales, transform=positive())
        self._validate_ard_active_dims(self.lengthscales)

    @property
    def ard(self) -> bool:  # stationaries.py:66-72
        return self.lengthscales.numpy().ndim > 0

    def _leaf_record(self, D: int) -> dict:
        rec = {"op": self._op, "variance": float(self.variance.numpy())}
        ls = self.lengthscales.numpy()
        if ls.ndim > 0:
            rec["ard"] = ls.reshape(-1)
        else:
            rec["lengthscale"] = float(ls)
        return rec


class IsotropicStationary(Stationary):
    pass


class SquaredExponential(IsotropicStationary):
    _op = _lib.K_RBF


class RationalQuadratic(IsotropicStationary):
    _op = _lib.K_RQ

    def __init__(self, variance: Any = 1.0, lengthscales: Any = 1.0, alpha: Any = 1.0,
                 active_dims: ActiveDims = None) -> None:
        super().__init__(variance=variance, lengthscales=lengthscales, active_dims=active_dims)
        self.alpha = Parameter(alpha, transform=positive())

    def _leaf_record(self, D: int) -> dict:
        rec = super()._leaf_record(D)
        rec["alpha"] = float(self.alpha.numpy())
        return rec


class Exponential(IsotropicStationary):
    _op = _lib.K_EXPONENTIAL


class Matern12(IsotropicStationary):
    _op = _lib.K_MATERN12


class Matern32(IsotropicStationary):
    _op = _lib.K_MATERN32


class Matern52(IsotropicStationary):
    _op = _lib.K_MATERN52
