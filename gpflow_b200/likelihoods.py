"""Gaussian likelihood (mirrors gpflow/likelihoods/scalar_continuous.py:41-148 for a constant
variance Parameter; heteroskedastic `Function` variances are outside the hot path)."""
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
            if callable(variance):
                raise NotImplementedError("heteroskedastic Gaussian(variance=Function) is outside the hot path")
            self.variance: Optional[Parameter] = Parameter(variance, transform=positive(lower=self.variance_lower_bound))
            self.scale: Optional[Parameter] = None
        else:
            assert variance is None, "Cannot set both `variance` and `scale`."
            self.variance = None
            self.scale = Parameter(scale, transform=positive(lower=self.scale_lower_bound))

    def _variance_value(self) -> float:  # scalar_continuous.py:92-102
        if self.variance is not None:
            return float(self.variance.numpy())
        return float(self.scale.numpy()) ** 2

    def variance_at(self, X):  # scalar_continuous.py:108-111 -> [N, 1]
        X = ops.to_device(X)
        return ops.full((X.shape[0], 1), self._variance_value(), like=X)

    def variational_expectations(self, X, Fmu, Fvar, Y):
        """Sum over the batch of scalar_continuous.py:139-148, returned as a device fp64 scalar [1].
        (The reference returns the per-row vector; every hot-path caller immediately reduce_sums it,
        svgp.py:181, so the reduction is fused.)"""
        return ops.gaussian_varexp_sum(ops.to_device(Fmu), ops.to_device(Fvar), ops.to_device(Y),
                                       self._variance_value())

    def predict_mean_and_var(self, X, Fmu, Fvar):  # scalar_continuous.py:127-130
        out = ops.copy(Fvar)
        ones = ops.full(Fvar.shape, self._variance_value(), like=Fvar)
        ops.axpby(1.0, ones, 1.0, out)
        return Fmu, out

    def predict_log_density(self, X, Fmu, Fvar, Y):  # scalar_continuous.py:133-136 -> device vector [N]
        Fmu, Fvar, Y = ops.to_device(Fmu), ops.to_device(Fvar), ops.to_device(Y)
        return ops.gaussian_log_density(Fmu, Fvar, Y, self._variance_value())
