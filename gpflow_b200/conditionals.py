"""Conditionals (mirrors gpflow/conditionals/util.py:37-169 and conditionals/conditionals.py:27-87)."""
from __future__ import annotations

from typing import Any, Optional, Tuple

from . import _lib, config, covariances, ops
from .inducing_variables import InducingVariables, inducingpoint_wrapper
from .kernels import Kernel


def base_conditional(Kmn, Kmm, Knn, f, *, full_cov: bool = False, q_sqrt=None, white: bool = False):
    """util.py:37-70: Lm = chol(Kmm), then base_conditional_with_lm."""
    Lm, dinv = ops.cholesky(ops.to_device(Kmm))
    return base_conditional_with_lm(Kmn, Lm, Knn, f, full_cov=full_cov, q_sqrt=q_sqrt, white=white, dinv=dinv)


def base_conditional_with_lm(Kmn, Lm, Knn, f, *, full_cov: bool = False, q_sqrt=None, white: bool = False,
                             dinv=None) -> Tuple[Any, Any]:
    """util.py:84-169 for Kmn [M, N] (no leading batch dims).
    Returns fmean [N, R] and fvar [N, R] (full_cov=False) or [R, N, N] (full_cov=True)."""
    Kmn, Lm, f = ops.to_device(Kmn), ops.to_device(Lm), ops.to_device(f)
    M, N = Kmn.shape
    R = f.shape[1]
    A = ops.trsm(Lm, ops.copy(Kmn), dinv=dinv)                        # util.py:125
    if full_cov:
        Knn = ops.to_device(Knn)
        base = ops.copy(Knn)
        ops.gemm(A, A, transa=True, alpha=-1.0, beta=1.0, out=base)   # util.py:129
    else:
        base = ops.copy(ops.to_device(Knn).reshape(-1))
        ops.colsumsq(A, scale=-1.0, out=base, accumulate=True)        # util.py:133
    if not white:
        ops.trsm(Lm, A, trans=True, dinv=dinv)                        # util.py:138-139
    fmean = ops.gemm(A, f, transa=True)                               # util.py:144
    if full_cov:
        fvar = ops.empty((R, N, N), like=A)
        for r in range(R):
            ops.axpby(1.0, base, 0.0, fvar[r])
    else:
        fvar_t = ops.empty((R, N), like=A)                            # [R, N], transposed at the end
        for r in range(R):
            ops.axpby(1.0, base, 0.0, fvar_t[r])
    if q_sqrt is not None:
        q_sqrt = ops.to_device(q_sqrt)
        if q_sqrt.dim() == 2:                                         # util.py:149 diagonal q_sqrt [M, R]
            for r in range(R):
                LTA = ops.copy(A)
                ops.scale_rows_(LTA, ops.copy(q_sqrt[:, r:r + 1]).reshape(-1))
                if full_cov:
                    ops.gemm(LTA, LTA, transa=True, beta=1.0, out=fvar[r])
                else:
                    ops.colsumsq(LTA, out=fvar_t[r], accumulate=True)
        elif q_sqrt.dim() == 3:                                       # util.py:151-157 lower-triangular [R, M, M]
            for r in range(R):
                if full_cov:
                    LTA = ops.gemm(q_sqrt[r], A, transa=True, flags=_lib.GPK_GEMM_A_LOWER)
                    ops.gemm(LTA, LTA, transa=True, beta=1.0, out=fvar[r])   # util.py:162
                else:                                                  # util.py:164, LTA never materialised
                    ops.gemm(q_sqrt[r], A, transa=True, out=fvar_t[r],
                             flags=_lib.GPK_GEMM_A_LOWER | _lib.GPK_GEMM_COLSUMSQ)
        else:
            raise ValueError("Bad dimension for q_sqrt: %s" % str(q_sqrt.dim()))
    if not full_cov:
        fvar = ops.transpose(fvar_t)                                  # util.py:167 -> [N, R]
    return fmean, fvar


def conditional(Xnew, inducing_variable, kernel: Kernel, f, *, full_cov: bool = False,
                full_output_cov: bool = False, q_sqrt=None, white: bool = False):
    """Single-output-kernel sparse conditional (conditionals.py:27-87): builds the posterior object and
    calls its fused prediction, exactly as the reference does."""
    from .posteriors import PrecomputeCacheType, create_posterior

    iv = inducingpoint_wrapper(inducing_variable)
    posterior = create_posterior(kernel, iv, f, q_sqrt, whiten=white, mean_function=None,
                                 precompute_cache=PrecomputeCacheType.NOCACHE)
    return posterior.fused_predict_f(Xnew, full_cov=full_cov, full_output_cov=full_output_cov)


def sample_mvn(mean, cov, full_cov: bool, num_samples: Optional[int] = None, *, eps=None, generator=None):
    """conditionals/util.py:179-211 for 2-D `mean` [N, D]: a sample (or `num_samples` samples, leading axis) from
    N(mean, cov) with cov [N, D] (full_cov=False: independent entries) or [N, D, D] (full_cov=True: one D x D covariance per
    row, factorised with the default jitter).  The standard-normal draws come from `eps` when given ([S, N, D] / [N, D, S], the
    reference's shapes) -- that is how the parity tests feed the oracle the same noise -- else from torch.randn (input
    generation; all arithmetic on the draws is libgpk)."""
    T = ops.torch()
    mean, cov = ops.to_device(mean), ops.to_device(cov)
    S = 1 if num_samples is None else int(num_samples)
    N, D = mean.shape
    if not full_cov:
        e = T.randn((S, N, D), dtype=mean.dtype, device=mean.device, generator=generator) if eps is None else ops.to_device(eps)
        sd = ops.copy(cov)
        _lib.check(_lib.load().gpk_clamp_min(ops._p(sd), N, D, ops._ld(sd), 0.0, 2, ops.dtype_code(sd), ops._stream()),
                   "gpk_clamp_min")
        out = ops.copy(e.reshape(S * N, D)).view(S, N, D)
        for s_ in range(S):
            ops.hadamard_(out[s_], sd)
            ops.axpby(1.0, mean, 1.0, out[s_])
    else:
        e = T.randn((N, D, S), dtype=mean.dtype, device=mean.device, generator=generator) if eps is None else ops.to_device(eps)
        tmp = ops.empty((N, S, D), like=mean)                                  # per row n: (chol_n eps_n)^T [S, D]
        for n in range(N):
            C = ops.copy(cov[n])
            ops.add_diag_(C, config.default_jitter())                          # util.py:202-204
            Lc, _ = ops.potrf(C)
            ops.tril_(Lc)
            smp = ops.gemm(Lc, e[n])                                           # [D, S]
            ops.transpose(smp, out=tmp[n])
        out = ops.empty((S, N, D), like=mean)
        for s_ in range(S):                                                    # [N, S, D] -> [S, N, D], + mean
            ops.axpby(1.0, tmp[:, s_, :], 0.0, out[s_])
            ops.axpby(1.0, mean, 1.0, out[s_])
    return out[0] if num_samples is None else out
