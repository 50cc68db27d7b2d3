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
        from .kernels import kernel_matrix

        if self.likelihood.heteroskedastic:  # add_noise_cov with the per-point variance (model_utils.py:33-50)
            s2, svec = 0.0, self.likelihood.variance_at(self.X_data).reshape(-1)
        else:
            s2, svec = self.likelihood._variance_value(), None
        Kmm = kernel_matrix(self.kernel, self.X_data, None, uplo=_lib.GPK_LOWER, diag_scalar=s2, diag_vec=svec)
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


class IndependentPosteriorMultiOutput(IndependentPosterior):
    """posteriors.py:844-885: L independent latent GPs; the kernel and / or the inducing variables may be shared.
    fmean [N, L]; fvar [N, L] (full_cov=False) or [L, N, N] (full_cov=True).  full_output_cov=True expands the
    independent outputs to (block-)diagonal form (conditionals/util.py:222-254)."""

    def _latents(self):
        return covariances._latent_pairs(self.X_data, self.kernel)

    def _conditional_fused(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        from .inducing_variables import FallbackSharedIndependentInducingVariables
        from .kernels import SharedIndependent

        q_mu, q_sqrt = self.q_mu, self.q_sqrt
        if isinstance(self.X_data, FallbackSharedIndependentInducingVariables) and isinstance(self.kernel, SharedIndependent):
            Knn = self.kernel.kernel(Xnew, full_cov=full_cov)                     # :851-860: one factorisation for all
            Kmm = covariances.Kuu(self.X_data, self.kernel, jitter=config.default_jitter())
            Kmn = covariances.Kuf(self.X_data, self.kernel, Xnew)
            Lm, dinv = ops.potrf(Kmm)
            ops.tril_(Lm)
            fmean, fvar = base_conditional_with_lm(Kmn, Lm, Knn, q_mu, full_cov=full_cov, q_sqrt=q_sqrt,
                                                   white=self.whiten, dinv=dinv)
        else:                                                                     # :861-882, one latent at a time
            pairs = self._latents()
            L, N = len(pairs), Xnew.shape[0]
            fmean_t = ops.empty((L, N), like=Xnew)
            fvar = ops.empty((L, N, N), like=Xnew) if full_cov else None
            fvar_t = None if full_cov else ops.empty((L, N), like=Xnew)
            for l, (iv, k) in enumerate(pairs):
                Kmm = covariances.Kuu(iv, k, jitter=config.default_jitter())
                Kmn = covariances.Kuf(iv, k, Xnew)
                Knn = k(Xnew, full_cov=full_cov)
                Lm, dinv = ops.potrf(Kmm)
                ops.tril_(Lm)
                qs = None if q_sqrt is None else (ops.copy(q_sqrt[:, l:l + 1]) if q_sqrt.dim() == 2 else q_sqrt[l:l + 1])
                m_l, v_l = base_conditional_with_lm(Kmn, Lm, Knn, ops.copy(q_mu[:, l:l + 1]), full_cov=full_cov, q_sqrt=qs,
                                                    white=self.whiten, dinv=dinv)
                ops.axpby(1.0, m_l.reshape(-1), 0.0, fmean_t[l])
                if full_cov:
                    ops.axpby(1.0, v_l[0], 0.0, fvar[l])
                else:
                    ops.axpby(1.0, v_l.reshape(-1), 0.0, fvar_t[l])
            fmean = ops.transpose(fmean_t)
            if not full_cov:
                fvar = ops.transpose(fvar_t)
        return self._post_process_mean_and_cov(fmean, fvar, full_cov, full_output_cov)

    def _conditional_with_precompute(self, cache, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        return self._conditional_fused(Xnew, full_cov=full_cov, full_output_cov=full_output_cov)

    def _precompute(self):
        return ()   # the per-latent factorisations are redone per call (cache type semantics kept: results identical)

    def _post_process_mean_and_cov(self, mean, cov, full_cov: bool, full_output_cov: bool):
        return mean, expand_independent_outputs(cov, full_cov, full_output_cov)


def expand_independent_outputs(fvar, full_cov: bool, full_output_cov: bool):
    """conditionals/util.py:222-254: [N, P] -> [N, P, P] / [P, N, N] -> [N, P, N, P] (diagonal over the outputs) when
    full_output_cov, identity otherwise."""
    if not full_output_cov:
        return fvar
    if full_cov:
        P, N = fvar.shape[0], fvar.shape[1]
        # block-diagonal in a [(p, n), (p', n')]-ordered buffer, then reordered to [n, p, n', p']
        buf = ops.full((P * N, P * N), 0.0, like=fvar)                 # rows (p, n), cols (p', n')
        for p in range(P):
            ops.axpby(1.0, fvar[p], 0.0, buf[p * N:(p + 1) * N, p * N:(p + 1) * N])
        return _pnpn_to_npnp(buf, P, N)
    N, P = fvar.shape
    out = ops.full((N, P * P), 0.0, like=fvar)
    for p in range(P):
        ops.axpby(1.0, fvar[:, p:p + 1], 0.0, out[:, p * P + p:p * P + p + 1])
    return out.view(N, P, P)


def _pnpn_to_npnp(buf, P: int, N: int):
    """[(p, n), (p', n')] -> [n, p, n', p'] with two blocked transpositions (rows, then columns)."""
    out = ops.empty((N * P, N * P), like=buf)
    tmp = ops.empty((N * P, N * P), like=buf)
    for p in range(P):          # rows: (p, n) -> (n, p)
        ops.axpby(1.0, buf[p * N:(p + 1) * N, :], 0.0, tmp[p::P, :])
    tmp_t = ops.transpose(tmp)  # now rows are (p', n')
    out_t = ops.empty((N * P, N * P), like=buf)
    for p in range(P):
        ops.axpby(1.0, tmp_t[p * N:(p + 1) * N, :], 0.0, out_t[p::P, :])
    return ops.transpose(out_t, out=out).view(N, P, N, P)


class LinearCoregionalizationPosterior(IndependentPosteriorMultiOutput):
    """posteriors.py:888-901: the latent posteriors mixed by W, f = W g (conditionals/util.py:518-563)."""

    def _post_process_mean_and_cov(self, mean, cov, full_cov: bool, full_output_cov: bool):
        import numpy as np

        W = np.asarray(self.kernel.W.numpy(), dtype=np.float64)              # [P, L]
        P, L = W.shape
        Wd = ops.to_device(W)
        f_mean = ops.gemm(mean, Wd, transb=True)                               # [N, P]
        if not full_cov and not full_output_cov:
            return f_mean, ops.gemm(cov, ops.to_device(W ** 2), transb=True)   # [N, L] x (W^2)^T
        if not full_cov and full_output_cov:                                   # [N, P, P] = sum_l g_var[n, l] W[:, l] W[:, l]^T
            outer = np.stack([np.outer(W[:, l], W[:, l]).reshape(-1) for l in range(L)])   # [L, P * P]
            return f_mean, ops.gemm(cov, ops.to_device(outer)).view(cov.shape[0], P, P)
        if full_cov and not full_output_cov:                                   # [P, N, N] = sum_l W[p, l]^2 g_var[l]
            N = cov.shape[1]
            out = ops.full((P, N, N), 0.0, like=cov)
            for p in range(P):
                for l in range(L):
                    ops.axpby(float(W[p, l] ** 2), cov[l], 1.0, out[p])
            return f_mean, out
        N = cov.shape[1]                                                       # [N, P, N, P]
        buf = ops.full((P * N, P * N), 0.0, like=cov)
        for p in range(P):
            for q in range(P):
                blk = buf[p * N:(p + 1) * N, q * N:(q + 1) * N]
                for l in range(L):
                    ops.axpby(float(W[p, l] * W[q, l]), cov[l], 1.0, blk)
        return f_mean, _pnpn_to_npnp(buf, P, N)


@get_posterior_class.register(MultioutputKernel, InducingVariables)
def _get_posterior_mo(kernel, inducing_variable):
    from .inducing_variables import (FallbackSeparateIndependentInducingVariables,
                                     FallbackSharedIndependentInducingVariables)
    from .kernels import LinearCoregionalization, SeparateIndependent, SharedIndependent

    if not isinstance(inducing_variable, (FallbackSharedIndependentInducingVariables,
                                          FallbackSeparateIndependentInducingVariables)):
        raise NotImplementedError("multi-output kernels need Shared/SeparateIndependentInducingVariables")
    if isinstance(kernel, LinearCoregionalization):
        return LinearCoregionalizationPosterior          # posteriors.py:1074-1086
    if isinstance(kernel, (SharedIndependent, SeparateIndependent)):
        return IndependentPosteriorMultiOutput           # posteriors.py:1050-1071
    raise NotImplementedError("fully correlated multi-output posteriors (posteriors.py:904-1009) are not built: every "
                              "kernel class of this package has independent latent GPs")


def create_posterior(kernel: Kernel, inducing_variable: InducingVariables, q_mu, q_sqrt, whiten: bool,
                     mean_function: Optional[MeanFunction] = None,
                     precompute_cache: Union[PrecomputeCacheType, str, None] = PrecomputeCacheType.TENSOR):
    """posteriors.py:1089-1108."""
    cls = get_posterior_class(kernel, inducing_variable)
    return cls(kernel, inducing_variable, q_mu, q_sqrt, whiten, mean_function,
               precompute_cache=_validate_precompute_cache_type(precompute_cache))
