"""Gaussian-process regression (mirrors gpflow/models/gpr.py:36-196)."""
from __future__ import annotations

import ctypes
from typing import Any, Optional

import numpy as np

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
        Yc = self._centred_targets()
        if self.likelihood.heteroskedastic:   # per-point noise (scalar_continuous.py:92-111; model_utils.py:33-50)
            s2, svec = 0.0, self.likelihood.variance_at(X).reshape(-1).contiguous()
        else:
            s2, svec = self.likelihood._variance_value(), None
        # a fresh result vector per call: earlier results stay valid when the model is evaluated again
        self._out = ops.torch().empty((4,), dtype=ops.torch().float64, device=X.device)
        if not self.kernel.is_fusable():
            return self._lml_unfused(X, Yc, s2, svec)
        need = lib.gpk_gpr_lml_ws(N, P, dc)
        if self._ws is None or self._ws.numel() < need:
            self._ws = ops.scratch_bytes(need)
        nodes, n_nodes, dims, ard = compile_kernel(self.kernel, D)
        _lib.check(lib.gpk_gpr_lml(nodes, n_nodes, dims, ard, ops._p(X), N, ops._ld(X), D, ops._p(Yc), P, s2, ops._p(svec),
                                   dc, ops._p(self._out), ops._p(self._ws), ops._stream()), "gpk_gpr_lml")
        return ops.objective(self._out, 0, 3)

    def _lml_unfused(self, X, Yc, s2, svec):
        """gpr.py:91-107 composed from the individual operators for kernels without a fused K-build record (Cosine,
        Periodic, ArcCosine, Coregion, ChangePoints and combinations with them): K materialised, (Y - m)^T riding along
        the factorisation as extra rows, the same reductions."""
        import math
        from ..kernels import kernel_matrix

        T = ops.torch()
        N, P = Yc.shape
        A = ops.empty((N + P, N), like=X)
        K = kernel_matrix(self.kernel, X, None, diag_scalar=s2, diag_vec=svec)
        ops.axpby(1.0, K, 0.0, A[:N])
        ops.transpose(Yc, out=A[N:])
        info = T.empty((1,), dtype=T.int32, device=X.device)
        lib = _lib.load()
        ws = ops.scratch_bytes(lib.gpk_potrf_ws(N, N + P, ops.dtype_code(A)))
        _lib.check(lib.gpk_potrf(ops._p(A), N, N + P, ops._ld(A), ops.dtype_code(A), ops._p(info), ops._p(ws),
                                 ops._stream()), "gpk_potrf")
        out = self._out
        ops.fill(out.view(1, 4), 0.0)
        ops.reduce(ops.SUMSQ, A[N:], N * P, 1, out=out[1:2], accumulate=True)            # sum alpha^2
        ops.reduce(ops.SUMLOG, A, N, ops._ld(A) + 1, out=out[2:3], accumulate=True)       # sum log diag L
        ops.axpby(-0.5, out[1:2], 0.0, out[0:1])
        ops.axpby(-float(P), out[2:3], 1.0, out[0:1])
        ops.axpby(1.0, ops.full((1,), -0.5 * N * P * math.log(2.0 * math.pi), dtype="float64"), 1.0, out[0:1])
        self._info = info
        return ops.objective(out, 0, None, info_tensor=info)

    def log_marginal_likelihood_and_grad(self):
        """Value and gradient in ONE fused call (gpk_gpr_lml_grad): the backward pass the reference gets from TensorFlow
        autodiff through gpr.py:91-107.  Returns (lml, grads): `lml` as log_marginal_likelihood(); `grads` a dict
        {Parameter: dLML/d(constrained value)} for the kernel variance, the lengthscale(s) and the likelihood variance
        (NumPy, after one small device->host read).  Covers a single stationary leaf kernel in float64."""
        from ..kernels.stationaries import Stationary

        k = self.kernel
        if not isinstance(k, Stationary) or k._op not in (_lib.K_RBF, _lib.K_MATERN12, _lib.K_MATERN32,
                                                          _lib.K_MATERN52, _lib.K_EXPONENTIAL):
            raise NotImplementedError("the device backward pass covers a single SquaredExponential / Matern12 / "
                                      "Matern32 / Matern52 / Exponential kernel")
        if self.likelihood.variance is None:
            raise NotImplementedError("the device backward pass covers Gaussian(variance=...)")
        lib = _lib.load()
        X, Y = self.data
        N, D = X.shape
        P = Y.shape[1]
        dc = ops.dtype_code(X)
        if dc != _lib.GPK_F64:
            raise NotImplementedError("the device backward pass computes in float64")
        need = lib.gpk_gpr_lml_grad_ws(N, P, dc)
        if getattr(self, "_gws", None) is None or self._gws.numel() < need:
            self._gws = ops.scratch_bytes(need)
        nl = int(k.lengthscales.numpy().size) if k.ard else 1
        out = ops.torch().empty((6 + nl,), dtype=ops.torch().float64, device=X.device)
        nodes, n_nodes, dims, ard = compile_kernel(k, D)
        Yc = self._centred_targets()
        _lib.check(lib.gpk_gpr_lml_grad(nodes, n_nodes, dims, ard, ops._p(X), N, ops._ld(X), D, ops._p(Yc), P,
                                        self.likelihood._variance_value(), dc, ops._p(out), 6 + nl, ops._p(self._gws),
                                        ops._stream()), "gpk_gpr_lml_grad")
        self._out = out
        h = out.cpu().numpy()
        if int(h[3]) != 0:
            raise ops.NonPositiveDefiniteError(f"Cholesky decomposition was not successful (pivot {int(h[3])} <= 0)")
        grads = {k.variance: np.asarray(h[4]), self.likelihood.variance: np.asarray(h[5]),
                 k.lengthscales: (h[6:6 + nl].copy() if k.ard else np.asarray(h[6]))}
        return ops.objective(out, 0, 3), grads

    def training_loss_and_gradients(self):
        """(loss, gradients) for the optimiser contract of gpflow/optimizers/scipy.py:322-331: loss = -LML (float) and one
        gradient per TRAINABLE parameter w.r.t. its UNCONSTRAINED variable, in `trainable_parameters` order."""
        lml, grads = self.log_marginal_likelihood_and_grad()
        out = []
        for p in self.trainable_parameters:
            if p not in grads:
                raise NotImplementedError("a trainable parameter has no device gradient (mean-function parameters, "
                                          "priors and data gradients are outside the hot path)")
            out.append(-p.unconstrained_gradient(grads[p]))
        return -float(lml), out

    def cholesky_info(self) -> int:
        """0, or the 1-based index of the first non-positive pivot of the last evaluation."""
        if getattr(self, "_info", None) is not None and not self.kernel.is_fusable():
            return int(self._info.item())
        return int(self._out[3].item()) if self._out is not None else 0

    def posterior(self, precompute_cache=posteriors.PrecomputeCacheType.TENSOR) -> posteriors.GPRPosterior:
        """gpr.py:146-175."""
        return posteriors.GPRPosterior(kernel=self.kernel, data=self.data, likelihood=self.likelihood,
                                       mean_function=self.mean_function,
                                       precompute_cache=posteriors._validate_precompute_cache_type(precompute_cache))

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):  # gpr.py:178-190
        return self.posterior(posteriors.PrecomputeCacheType.NOCACHE).fused_predict_f(
            Xnew, full_cov=full_cov, full_output_cov=full_output_cov)
