"""Posterior objects with fused and cached prediction (mirrors gpflow/posteriors.py:97-169, 193-358,
361-562, 639-841, 1039-1108)."""
from __future__ import annotations

import enum
from abc import ABC, abstractmethod
from typing import Any, Optional, Tuple, Type, Union

from . import _lib, config, covariances, ops
from .base import Module
from .conditionals import base_conditional, base_conditional_with_lm
from .inducing_variables import InducingPoints, InducingVariables
from .kernels import Kernel, MultioutputKernel
from .likelihoods import Gaussian
from .mean_functions import MeanFunction, Zero
from .utilities.multipledispatch import Dispatcher


class PrecomputeCacheType(enum.Enum):
    """posteriors.py:97-114."""

    TENSOR = "tensor"
    VARIABLE = "variable"
    NOCACHE = "nocache"


def _validate_precompute_cache_type(value) -> PrecomputeCacheType:  # posteriors.py:172-190
    if value is None:
        return PrecomputeCacheType.NOCACHE
    if isinstance(value, PrecomputeCacheType):
        return value
    if isinstance(value, str):
        return PrecomputeCacheType(value.lower())
    raise ValueError(f"{value} is not a valid PrecomputeCacheType. Valid options: 'tensor', 'variable', "
                     "'nocache' (or None).")


def _assert_params_false(called: str, **kwargs: bool) -> None:
    """gpflow/utilities/model_utils.py:10-25."""
    errors = ", ".join(f"{k}={v}" for k, v in kwargs.items() if v)
    if errors:
        raise NotImplementedError(f"{called} does not currently support: {errors}")


class AbstractPosterior(Module, ABC):
    def __init__(self, kernel: Kernel, X_data, cache: Optional[Tuple[Any, ...]] = None,
                 mean_function: Optional[MeanFunction] = None) -> None:
        self.kernel = kernel
        self.X_data = X_data
        self.cache = cache
        self.mean_function = mean_function
        self._precompute_cache: Optional[PrecomputeCacheType] = None

    def _add_mean_function(self, Xnew, mean):  # posteriors.py:225-229
        if self.mean_function is None or isinstance(self.mean_function, Zero):
            return mean
        ops.axpby(1.0, self.mean_function(Xnew), 1.0, mean)
        return mean

    @abstractmethod
    def _precompute(self) -> Tuple[Any, ...]:
        ...

    def fused_predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):  # :248-258
        Xnew = ops.to_device(Xnew)
        mean, cov = self._conditional_fused(Xnew, full_cov=full_cov, full_output_cov=full_output_cov)
        return self._add_mean_function(Xnew, mean), cov

    @abstractmethod
    def _conditional_fused(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        ...

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):  # :285-299
        if self.cache is None:
            raise ValueError("Cache has not been precomputed yet. Call update_cache first or use fused_predict_f")
        Xnew = ops.to_device(Xnew)
        mean, cov = self._conditional_with_precompute(self.cache, Xnew, full_cov=full_cov,
                                                      full_output_cov=full_output_cov)
        return self._add_mean_function(Xnew, mean), cov

    @abstractmethod
    def _conditional_with_precompute(self, cache, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        ...

    def update_cache(self, precompute_cache: Optional[PrecomputeCacheType] = None) -> None:  # :322-358
        if precompute_cache is None:
            if self._precompute_cache is None:
                raise ValueError("You must pass precompute_cache explicitly (the cache had not been updated before).")
            precompute_cache = self._precompute_cache
        else:
            self._precompute_cache = precompute_cache
        if precompute_cache is PrecomputeCacheType.NOCACHE:
            self.cache = None
        else:  # TENSOR and VARIABLE coincide here: device buffers, refreshed in place by the next update
            self.cache = tuple(self._precompute())


class GPRPosterior(AbstractPosterior):
    """posteriors.py:361-443."""

    def __init__(self, kernel: Kernel, data, likelihood: Gaussian, mean_function: MeanFunction, *,
                 precompute_cache: Optional[PrecomputeCacheType]) -> None:
        X, Y = data
        super().__init__(kernel, ops.to_device(X), mean_function=mean_function)
        self.Y_data = ops.to_device(Y)
        self.likelihood = likelihood
        if precompute_cache is not None:
            self.update_cache(precompute_cache)

    def _precompute(self):  # :415-432 — err, Lm = chol(K + sigma^2 I), no jitter
        err = ops.copy(self.Y_data)
        if self.mean_function is not None and not isinstance(self.mean_function, Zero):
            ops.axpby(-1.0, self.mean_function(self.X_data), 1.0, err)
        from .kernels import compile_kernel

        desc = compile_kernel(self.kernel, self.X_data.shape[1])
        Kmm = ops.kbuild(desc, self.X_data, None, uplo=_lib.GPK_LOWER,
                         diag_scalar=self.likelihood._variance_value())
        Lm, dinv = ops.potrf(Kmm)
        ops.tril_(Lm)
        return err, Lm, dinv

    def _conditional_with_precompute(self, cache, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        _assert_params_false("GPRPosterior._conditional_with_precompute", full_output_cov=full_output_cov)
        err, Lm, dinv = cache
        Knn = self.kernel(Xnew, full_cov=full_cov)   # :402
        Kmn = self.kernel(self.X_data, Xnew)         # :403
        return base_conditional_with_lm(Kmn, Lm, Knn, err, full_cov=full_cov, q_sqrt=None, white=False, dinv=dinv)

    def _conditional_fused(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):  # :435-443
        return self._conditional_with_precompute(tuple(self._precompute()), Xnew, full_cov, full_output_cov)


class SGPRPosterior(AbstractPosterior):
    """posteriors.py:446-562."""

    def __init__(self, kernel: Kernel, data, inducing_variable: InducingPoints, likelihood: Gaussian,
                 num_latent_gps: int, mean_function: MeanFunction, *,
                 precompute_cache: Optional[PrecomputeCacheType]) -> None:
        X, Y = data
        super().__init__(kernel, ops.to_device(X), mean_function=mean_function)
        self.Y_data = ops.to_device(Y)
        self.likelihood = likelihood
        self.inducing_variable = inducing_variable
        self.num_latent_gps = num_latent_gps
        if precompute_cache is not None:
            self.update_cache(precompute_cache)

    def _precompute(self):  # :520-551 via the fused SGPR evaluation (L, LB, c fall out of the ELBO pass)
        from .models.sgpr import _sgpr_fused

        M = self.inducing_variable.num_inducing
        P = self.Y_data.shape[1]
        L = ops.empty((M, M), like=self.X_data)
        LB = ops.empty((M, M), like=self.X_data)
        c = ops.empty((M, P), like=self.X_data)
        _sgpr_fused(self.X_data, self.Y_data, self.kernel, self.inducing_variable, self.likelihood,
                    self.mean_function, cache=(L, LB, c), owner=self)
        return L, LB, c

    def _conditional_with_precompute(self, cache, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        _assert_params_false("SGPRPosterior._conditional_with_precompute", full_output_cov=full_output_cov)
        L, LB, c = cache
        P = self.num_latent_gps
        Kus = covariances.Kuf(self.inducing_variable, self.kernel, Xnew)   # :495
        tmp1 = ops.trsm(L, Kus)                                            # :496
        Nn = Xnew.shape[0]
        if full_cov:
            var = ops.copy(self.kernel(Xnew))
            ops.gemm(tmp1, tmp1, transa=True, alpha=-1.0, beta=1.0, out=var)
        else:
            v = self.kernel(Xnew, full_cov=False)
            ops.colsumsq(tmp1, scale=-1.0, out=v, accumulate=True)
        tmp2 = ops.trsm(LB, tmp1)                                          # :497 (tmp1 consumed above)
        mean = ops.gemm(tmp2, c, transa=True)                              # :498
        if full_cov:
            ops.gemm(tmp2, tmp2, transa=True, alpha=1.0, beta=1.0, out=var)
            out = ops.empty((P, Nn, Nn), like=var)
            for p in range(P):                                             # :504 tile
                ops.axpby(1.0, var, 0.0, out[p])
            return mean, out
        ops.colsumsq(tmp2, out=v, accumulate=True)
        out_t = ops.empty((P, Nn), like=v)
        for p in range(P):                                                 # :511 tile
            ops.axpby(1.0, v, 0.0, out_t[p])
        return mean, ops.transpose(out_t)

    def _conditional_fused(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):  # :553-562
        return self._conditional_with_precompute(tuple(self._precompute()), Xnew, full_cov, full_output_cov)


class BasePosterior(AbstractPosterior):
    """posteriors.py:639-746 — q(u) = N(q_mu, q_sqrt q_sqrt^T) posteriors; cache = (alpha, Qinv)."""

    def __init__(self, kernel: Kernel, inducing_variable: InducingVariables, q_mu, q_sqrt, whiten: bool = True,
                 mean_function: Optional[MeanFunction] = None, *, precompute_cache: Optional[PrecomputeCacheType]):
        super().__init__(kernel, inducing_variable, mean_function=mean_function)
        self.whiten = whiten
        self._q_mu_src, self._q_sqrt_src = q_mu, q_sqrt
        if precompute_cache is not None:
            self.update_cache(precompute_cache)

    @property
    def q_mu(self):
        return ops.to_device(self._q_mu_src)

    @property
    def q_sqrt(self):
        return None if self._q_sqrt_src is None else ops.to_device(self._q_sqrt_src)

    def _precompute(self):  # :694-746 (single-output kernel: Kuu [M, M])
        Kuu = covariances.Kuu(self.X_data, self.kernel, jitter=config.default_jitter())
        q_mu, q_sqrt = self.q_mu, self.q_sqrt
        M, P = q_mu.shape
        L, dinv = ops.potrf(Kuu)
        ops.tril_(L)
        alpha = ops.copy(q_mu)
        if not self.whiten:
            ops.trsm(L, alpha, dinv=dinv)                                 # cholesky_solve :708
        ops.trsm(L, alpha, trans=True, dinv=dinv)                         # :708 / :710
        Qinv = ops.empty((P, M, M), like=q_mu)
        for p in range(P):
            B = ops.add_diag_(ops.full((M, M), 0.0, like=q_mu), 1.0)      # I
            if q_sqrt is not None:
                if q_sqrt.dim() == 2:
                    qs = ops.full((M, M), 0.0, like=q_mu)
                    ops.add_diag_(qs, 0.0, ops.copy(q_sqrt[:, p:p + 1]).reshape(-1))
                else:
                    qs = ops.tril_(ops.copy(q_sqrt[p]))
                if not self.whiten:
                    ops.trsm(L, qs, dinv=dinv)                            # :727
                ops.gemm(qs, qs, transb=True, alpha=-1.0, beta=1.0, out=B)  # B = I - C  (:728-737)
            ops.trsm(L, B, trans=True, dinv=dinv)                         # LinvT_B  :739
            Bt = ops.transpose(B)                                         # B_Linv   :740
            ops.trsm(L, Bt, trans=True, dinv=dinv)                        # Qinv     :741
            ops.axpby(1.0, Bt, 0.0, Qinv[p])
        return alpha, Qinv


class IndependentPosterior(BasePosterior):
    def _get_Kff(self, Xnew, full_cov: bool):  # :775-792 (single-output branch)
        return self.kernel(Xnew, full_cov=full_cov)

    def _conditional_with_precompute(self, cache, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        _assert_params_false("IndependentPosterior._conditional_with_precompute", full_output_cov=full_output_cov)
        alpha, Qinv = cache                                                 # :802-822
        Kuf = covariances.Kuf(self.X_data, self.kernel, Xnew)              # [M, N]
        Kff = self._get_Kff(Xnew, full_cov)
        mean = ops.gemm(Kuf, alpha, transa=True)                           # :808
        P, N = Qinv.shape[0], Xnew.shape[0]
        if full_cov:
            cov = ops.empty((P, N, N), like=Kuf)
            for p in range(P):
                QK = ops.gemm(Qinv[p], Kuf)
                ops.axpby(1.0, Kff, 0.0, cov[p])
                ops.gemm(Kuf, QK, transa=True, alpha=-1.0, beta=1.0, out=cov[p])   # :813-814
            return mean, cov
        cov_t = ops.empty((P, N), like=Kuf)
        for p in range(P):                                                  # :818-819: sum_m Kuf * (Qinv Kuf)
            # Qinv is symmetric: Kuf^T Qinv Kuf diag = colsum(Kuf * (Qinv Kuf)); with Qinv = R^T S R form
            QK = ops.gemm(Qinv[p], Kuf)
            prod = _colsum_prod(Kuf, QK)
            ops.axpby(1.0, Kff, 0.0, cov_t[p])
            ops.axpby(-1.0, prod, 1.0, cov_t[p])
        return mean, ops.transpose(cov_t)


def _colsum_prod(A, B):
    """sum_m A[m, n] * B[m, n] via the polarisation identity on fused column-sum-of-squares kernels:
    sum A*B = ( sum (A+B)^2 - sum (A-B)^2 ) / 4."""
    S = ops.copy(A)
    ops.axpby(1.0, B, 1.0, S)
    D = ops.copy(A)
    ops.axpby(-1.0, B, 1.0, D)
    out = ops.colsumsq(S, scale=0.25)
    ops.colsumsq(D, scale=-0.25, out=out, accumulate=True)
    return out


class IndependentPosteriorSingleOutput(IndependentPosterior):
    def _conditional_fused(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):  # :827-841
        _assert_params_false("IndependentPosteriorSingleOutput._conditional_fused", full_output_cov=full_output_cov)
        Knn = self.kernel(Xnew, full_cov=full_cov)
        Kmm = covariances.Kuu(self.X_data, self.kernel, jitter=config.default_jitter())
        Kmn = covariances.Kuf(self.X_data, self.kernel, Xnew)
        Lm, dinv = ops.potrf(Kmm)  # Kmm is a fresh buffer: factor in place (util.py:67)
        ops.tril_(Lm)
        return base_conditional_with_lm(Kmn, Lm, Knn, self.q_mu, full_cov=full_cov, q_sqrt=self.q_sqrt,
                                        white=self.whiten, dinv=dinv)


get_posterior_class = Dispatcher("get_posterior_class")


@get_posterior_class.register(Kernel, InducingVariables)
def _get_posterior_base_case(kernel: Kernel, inducing_variable: InducingVariables) -> Type[BasePosterior]:
    return IndependentPosteriorSingleOutput  # posteriors.py:1042-1047


@get_posterior_class.register(MultioutputKernel, InducingVariables)
def _get_posterior_mo(kernel, inducing_variable):
    raise NotImplementedError("multi-output SVGP posteriors (posteriors.py:844-1036) are outside the hot path; "
                              "shard independent outputs over models instead")


def create_posterior(kernel: Kernel, inducing_variable: InducingVariables, q_mu, q_sqrt, whiten: bool,
                     mean_function: Optional[MeanFunction] = None,
                     precompute_cache: Union[PrecomputeCacheType, str, None] = PrecomputeCacheType.TENSOR):
    """posteriors.py:1089-1108."""
    cls = get_posterior_class(kernel, inducing_variable)
    return cls(kernel, inducing_variable, q_mu, q_sqrt, whiten, mean_function,
               precompute_cache=_validate_precompute_cache_type(precompute_cache))
