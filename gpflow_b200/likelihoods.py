"""Gaussian likelihood (mirrors gpflow/likelihoods/scalar_continuous.py:41-148): a constant variance / scale
Parameter, or heteroskedastic -- `variance=Function` / `scale=Function` (any callable Module mapping X [N, D] to a device
tensor [N, 1], e.g. the mean functions) clipped from below as in the reference."""
from __future__ import annotations

import math
from typing import Any, Optional

import numpy as np

from . import config, ops
from .base import Module, Parameter, positive


class Likelihood(Module):
    pass


class ScalarLikelihood(Likelihood):
    pass


class Gaussian(ScalarLikelihood):
    def __init__(self, variance: Any = None, *, scale: Any = None, variance_lower_bound: Optional[float] = None):
        self.variance_lower_bound = (config.default_likelihood_positive_minimum()
                                     if variance_lower_bound is None else variance_lower_bound)
        self.scale_lower_bound = math.sqrt(self.variance_lower_bound)
        if scale is None:
            if variance is None:
                variance = 1.0
            # prepare_parameter_or_function (likelihoods/utils.py): a Function is kept, a constant becomes a Parameter
            self.variance: Any = variance if callable(variance) else Parameter(
                variance, transform=positive(lower=self.variance_lower_bound))
            self.scale: Any = None
        else:
            assert variance is None, "Cannot set both `variance` and `scale`."
            self.variance = None
            self.scale = scale if callable(scale) else Parameter(scale, transform=positive(lower=self.scale_lower_bound))

    @property
    def heteroskedastic(self) -> bool:
        return callable(self.variance) or callable(self.scale)

    def _variance_value(self) -> float:  # scalar_continuous.py:92-102, constant case
        if self.heteroskedastic:
            raise NotImplementedError("this operator takes a constant noise variance; heteroskedastic Gaussian "
                                      "likelihoods go through variance_at(X) (GPR, predict_y, predict_log_density)")
        if self.variance is not None:
            return float(self.variance.numpy())
        return float(self.scale.numpy()) ** 2

    def variance_at(self, X):  # scalar_continuous.py:92-111 -> device [N, 1]
        X = ops.to_device(X)
        if not self.heteroskedastic:
            return ops.full((X.shape[0], 1), self._variance_value(), like=X)
        fn, lower, square = (self.variance, self.variance_lower_bound, 0) if self.variance is not None else (
            self.scale, self.scale_lower_bound, 1)
        v = ops.copy(ops.to_device(fn(X)))
        if v.dim() != 2 or v.shape[0] != X.shape[0] or v.shape[1] != 1:
            raise ValueError(f"the noise Function must return [N, 1], got {tuple(v.shape)}")
        from . import _lib
        _lib.check(_lib.load().gpk_clamp_min(ops._p(v), v.shape[0], 1, 1, float(lower), square, ops.dtype_code(v),
                                             ops._stream()), "gpk_clamp_min")
        return v

    def variational_expectations(self, X, Fmu, Fvar, Y):
        """Sum over the batch of scalar_continuous.py:139-148, returned as a device fp64 scalar [1].
        (The reference returns the per-row vector; every hot-path caller immediately reduce_sums it,
        svgp.py:181, so the reduction is fused.)"""
        return ops.gaussian_varexp_sum(ops.to_device(Fmu), ops.to_device(Fvar), ops.to_device(Y),
                                       self._variance_value())

    def predict_mean_and_var(self, X, Fmu, Fvar):  # scalar_continuous.py:127-130
        out = ops.copy(Fvar)
        ops.axpby(1.0, self.variance_at(X), 1.0, out)   # [N, 1] broadcasts over the output columns
        return Fmu, out

    def predict_log_density(self, X, Fmu, Fvar, Y):  # scalar_continuous.py:133-136 -> device vector [N]
        Fmu, Fvar, Y = ops.to_device(Fmu), ops.to_device(Fvar), ops.to_device(Y)
        if self.heteroskedastic:
            tot = ops.copy(Fvar)
            ops.axpby(1.0, self.variance_at(X), 1.0, tot)
            return ops.gaussian_log_density(Fmu, tot, Y, 0.0)
        return ops.gaussian_log_density(Fmu, Fvar, Y, self._variance_value())
