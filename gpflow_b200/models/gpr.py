"""Gaussian-process regression (mirrors gpflow/models/gpr.py:36-196)."""
from __future__ import annotations

import ctypes
from typing import Any, Optional

from .. import _lib, ops, posteriors
from ..kernels import Kernel, compile_kernel
from ..likelihoods import Gaussian
from ..mean_functions import MeanFunction, Zero
from .model import GPModel, InternalDataTrainingLossMixin, data_input_to_tensor


class GPR(GPModel, InternalDataTrainingLossMixin):
    def __init__(self, data, kernel: Kernel, mean_function: Optional[MeanFunction] = None,
                 noise_variance: Any = None, likelihood: Optional[Gaussian] = None):
        assert (noise_variance is None) or (likelihood is None), "Cannot set both `noise_variance` and `likelihood`."
        if likelihood is None:
            if noise_variance is None:
                noise_variance = 1.0  # gpr.py:75-78
            likelihood = Gaussian(noise_variance)
        _, Y_data = data
        super().__init__(kernel, likelihood, mean_function, num_latent_gps=Y_data.shape[-1])
        self.data = data_input_to_tensor(data)
        self._ws = None
        self._out = None

    def maximum_log_likelihood_objective(self):  # gpr.py:85-86
        return self.log_marginal_likelihood()

    def _centred_targets(self):
        X, Y = self.data
        if isinstance(self.mean_function, Zero):
            return Y
        Yc = ops.copy(Y)
        return ops.axpby(-1.0, self.mean_function(X), 1.0, Yc)

    def log_marginal_likelihood(self):
        """gpr.py:91-107 in ONE fused call (gpk_gpr_lml): lower-triangle K-build with the noise on the
        diagonal, blocked Cholesky with (Y-m)^T riding along as extra rows, log-density reduction.
        Returns a device fp64 tensor of shape [] (float() it to synchronise)."""
        lib = _lib.load()
        X, Y = self.data
        N, D = X.shape
        P = Y.shape[1]
        dc = ops.dtype_code(X)
        need = lib.gpk_gpr_lml_ws(N, P, dc)
        if self._ws is None or self._ws.numel() < need:
            self._ws = ops.scratch_bytes(need)
            self._out = ops.torch().empty((4,), dtype=ops.torch().float64, device=X.device)
        nodes, n_nodes, dims, ard = compile_kernel(self.kernel, D)
        Yc = self._centred_targets()
        _lib.check(lib.gpk_gpr_lml(nodes, n_nodes, dims, ard, ops._p(X), N, ops._ld(X), D, ops._p(Yc), P,
                                   self.likelihood._variance_value(), None, dc, ops._p(self._out), ops._p(self._ws),
                                   ops._stream()), "gpk_gpr_lml")
        return self._out[0]

    def cholesky_info(self) -> int:
        """0, or the 1-based index of the first non-positive pivot of the last evaluation."""
        return int(self._out[3].item()) if self._out is not None else 0

    def posterior(self, precompute_cache=posteriors.PrecomputeCacheType.TENSOR) -> posteriors.GPRPosterior:
        """gpr.py:146-175."""
        return posteriors.GPRPosterior(kernel=self.kernel, data=self.data, likelihood=self.likelihood,
                                       mean_function=self.mean_function,
                                       precompute_cache=posteriors._validate_precompute_cache_type(precompute_cache))

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):  # gpr.py:178-190
        return self.posterior(posteriors.PrecomputeCacheType.NOCACHE).fused_predict_f(
            Xnew, full_cov=full_cov, full_output_cov=full_output_cov)
