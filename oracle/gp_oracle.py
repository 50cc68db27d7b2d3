"""CPU oracle for the GP-inference hot path (TEST INFRASTRUCTURE ONLY).

This file is a NumPy/SciPy restatement of the reference's algorithm for the path
K-build -> Cholesky/TRSM -> GPR LML / SGPR ELBO / SVGP ELBO / posterior mean+var.
Every function cites the reference file:line (relative to /root/reference, GPflow
2.9.2) whose operation order it follows.

It is the CHECKER, never the product: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  Nothing under
`gpflow_b200/` imports it, and the product path raises if the CUDA library is missing.

How parity is pinned.  The reference is pure Python on TensorFlow; TensorFlow,
TensorFlow-Probability, check_shapes and multipledispatch are NOT installed in this
image and there is no network, so the reference cannot be imported or executed here
(SURVEY.md section 8c).  The reference's own tests hold no golden LML/ELBO numbers;
they pin results through known-answer restatements and identities.  This oracle is
pinned against exactly those (tests/test_oracle_pins.py):
  * tests/gpflow/kernels/reference.py:13-27   O(N^2)-loop RBF (`ref_rbf_kernel`)
  * tests/gpflow/test_logdensities.py:113-128 scipy.stats.multivariate_normal.logpdf
  * tests/gpflow/test_kullback_leiblers.py:94-98,215-229 closed-form 1-D KL
  * tests/gpflow/test_kullback_leiblers.py:122-191 K vs K_cholesky / diag vs dense / white
  * tests/gpflow/conditionals/test_conditionals.py:168-214 explicit-inverse conditional
  * tests/integration/test_method_equivalence.py:181-241 SGPR(Z=X) == GPR, SVGP(q*) == SGPR
  * tests/gpflow/models/test_sgpr.py:29-44 compute_qu == predict_f(Z)
Outputs of the reference itself are unavailable; that limit is stated in DESIGN.md.

dtype: every function computes in the dtype of its inputs (fp64 or fp32), like the
reference under `default_float()` (gpflow/base.py:299-311).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.linalg as sla

LOG2PI = math.log(2.0 * math.pi)
DEFAULT_JITTER = 1e-6  # gpflow/config/__config__.py:98-104


# ----------------------------------------------------------------------------------------
# kernels
# ----------------------------------------------------------------------------------------
def square_distance(X: np.ndarray, X2: Optional[np.ndarray]) -> np.ndarray:
    """gpflow/utilities/ops.py:105-122 — norm expansion, may go slightly negative."""
    if X2 is None:
        Xs = np.sum(np.square(X), axis=-1, keepdims=True)
        dist = -2 * (X @ X.T)
        dist += Xs + Xs.T
        return dist
    Xs = np.sum(np.square(X), axis=-1)
    X2s = np.sum(np.square(X2), axis=-1)
    dist = -2 * (X @ X2.T)
    dist += Xs[:, None] + X2s[None, :]
    return dist


ActiveDims = Union[None, slice, Sequence[int]]


class Kernel:
    """gpflow/kernels/base.py:29-220 (active_dims slicing, __call__ routing, + and *)."""

    def __init__(self, active_dims: ActiveDims = None):
        if active_dims is None:
            active_dims = slice(None, None, None)
        elif not isinstance(active_dims, slice):
            active_dims = np.array(active_dims, dtype=int)
        self.active_dims = active_dims

    def slice(self, X, X2=None):  # base.py:90-109
        dims = self.active_dims
        X = X[..., dims]
        if X2 is not None:
            X2 = X2[..., dims]
        return X, X2

    def K(self, X, X2=None):
        raise NotImplementedError

    def K_diag(self, X):
        raise NotImplementedError

    def __call__(self, X, X2=None, *, full_cov=True, presliced=False):  # base.py:195-214
        if (not full_cov) and (X2 is not None):
            raise ValueError("Ambiguous inputs: `not full_cov` and `X2` are not compatible.")
        if not presliced:
            X, X2 = self.slice(X, X2)
        if not full_cov:
            return self.K_diag(X)
        return self.K(X, X2)

    def __add__(self, other):
        return Sum([self, other])

    def __mul__(self, other):
        return Product([self, other])


class Stationary(Kernel):
    """gpflow/kernels/stationaries.py:35-130."""

    def __init__(self, variance=1.0, lengthscales=1.0, active_dims=None):
        super().__init__(active_dims)
        self.variance = variance
        self.lengthscales = lengthscales

    def _c(self, v, like):
        return np.asarray(v, dtype=like.dtype)

    def scale(self, X):  # stationaries.py:77-79
        return X / self._c(self.lengthscales, X) if X is not None else X

    def K_diag(self, X):  # stationaries.py:82-83
        return np.full(X.shape[:-1], self.variance, dtype=X.dtype)

    def scaled_squared_euclid_dist(self, X, X2=None):  # stationaries.py:124-130
        return square_distance(self.scale(X), self.scale(X2))

    def K(self, X, X2=None):  # stationaries.py:103-105
        return self.K_r2(self.scaled_squared_euclid_dist(X, X2))

    def K_r2(self, r2):  # stationaries.py:111-116
        r = np.sqrt(np.maximum(r2, np.asarray(1e-36, dtype=r2.dtype)))
        return self.K_r(r)


class SquaredExponential(Stationary):
    def K_r2(self, r2):  # stationaries.py:209-210
        return self._c(self.variance, r2) * np.exp(-0.5 * r2)


RBF = SquaredExponential


class RationalQuadratic(Stationary):
    def __init__(self, variance=1.0, lengthscales=1.0, alpha=1.0, active_dims=None):
        super().__init__(variance, lengthscales, active_dims)
        self.alpha = alpha

    def K_r2(self, r2):  # stationaries.py:237-238
        a = self._c(self.alpha, r2)
        return self._c(self.variance, r2) * (1 + 0.5 * r2 / a) ** (-a)


class Exponential(Stationary):
    def K_r(self, r):  # stationaries.py:250-251
        return self._c(self.variance, r) * np.exp(-0.5 * r)


class Matern12(Stationary):
    def K_r(self, r):  # stationaries.py:270-271
        return self._c(self.variance, r) * np.exp(-r)


class Matern32(Stationary):
    def K_r(self, r):  # stationaries.py:290-292
        sqrt3 = self._c(np.sqrt(3.0), r)
        return self._c(self.variance, r) * (1.0 + sqrt3 * r) * np.exp(-sqrt3 * r)


class Matern52(Stationary):
    def K_r(self, r):  # stationaries.py:311-313
        sqrt5 = self._c(np.sqrt(5.0), r)
        c53 = self._c(5.0 / 3.0, r)
        return self._c(self.variance, r) * (1.0 + sqrt5 * r + c53 * np.square(r)) * np.exp(-sqrt5 * r)


class Static(Kernel):
    def __init__(self, variance=1.0, active_dims=None):
        super().__init__(active_dims)
        self.variance = variance

    def K_diag(self, X):  # statics.py:41-42
        return np.full(X.shape[:-1], self.variance, dtype=X.dtype)


class White(Static):
    def K(self, X, X2=None):  # statics.py:57-63 — zeros whenever X2 is given
        if X2 is None:
            return np.diag(np.full(X.shape[0], self.variance, dtype=X.dtype))
        return np.zeros((X.shape[0], X2.shape[0]), dtype=X.dtype)


class Constant(Static):
    def K(self, X, X2=None):  # statics.py:78-91
        n2 = X.shape[0] if X2 is None else X2.shape[0]
        return np.full((X.shape[0], n2), self.variance, dtype=X.dtype)


class Linear(Kernel):
    """gpflow/kernels/linears.py:25-68."""

    def __init__(self, variance=1.0, active_dims=None):
        super().__init__(active_dims)
        self.variance = variance

    def K(self, X, X2=None):  # linears.py:60-64
        v = np.asarray(self.variance, dtype=X.dtype)
        if X2 is None:
            return (X * v) @ X.T
        return (X * v) @ X2.T

    def K_diag(self, X):  # linears.py:67-68
        return np.sum(np.square(X) * np.asarray(self.variance, dtype=X.dtype), axis=-1)


class Polynomial(Linear):
    """gpflow/kernels/linears.py:71-112: (sigma^2 x.y + offset)^degree."""

    def __init__(self, degree=3.0, variance=1.0, offset=1.0, active_dims=None):
        super().__init__(variance, active_dims)
        self.degree, self.offset = degree, offset

    def K(self, X, X2=None):  # linears.py:107-108
        return (super().K(X, X2) + np.asarray(self.offset, dtype=X.dtype)) ** np.asarray(self.degree, dtype=X.dtype)

    def K_diag(self, X):  # linears.py:111-112
        return (super().K_diag(X) + np.asarray(self.offset, dtype=X.dtype)) ** np.asarray(self.degree, dtype=X.dtype)


class Combination(Kernel):
    """gpflow/kernels/base.py:223-302 — flattens same-class nesting; children see unsliced X."""

    def __init__(self, kernels: Sequence[Kernel]):
        super().__init__(None)
        flat: List[Kernel] = []
        for k in kernels:
            if isinstance(k, self.__class__):
                flat.extend(k.kernels)
            else:
                flat.append(k)
        self.kernels = flat

    def __call__(self, X, X2=None, *, full_cov=True, presliced=False):  # base.py:281-291
        return self._reduce([k(X, X2, full_cov=full_cov, presliced=presliced) for k in self.kernels])

    def K(self, X, X2=None):
        return self._reduce([k.K(X, X2) for k in self.kernels])

    def K_diag(self, X):
        return self._reduce([k.K_diag(X) for k in self.kernels])


class Sum(Combination):
    def _reduce(self, mats):  # base.py:305-308 (tf.add_n)
        out = mats[0]
        for m in mats[1:]:
            out = out + m
        return out


class Product(Combination):
    def _reduce(self, mats):  # base.py:311-314 (reduce(tf.multiply))
        out = mats[0]
        for m in mats[1:]:
            out = out * m
        return out


def difference_matrix(X, X2):
    """gpflow/utilities/ops.py:125-160 -> [N, N2, D]."""
    X2 = X if X2 is None else X2
    return X[:, None, :] - X2[None, :, :]


class Cosine(Stationary):
    """gpflow/kernels/stationaries.py:316-332 (AnisotropicStationary: K_d of the scaled per-dimension differences,
    stationaries.py:133-196)."""

    def K(self, X, X2=None):
        d = difference_matrix(self.scale(X), self.scale(X2))
        return self._c(self.variance, X) * np.cos(2 * np.pi * np.sum(d, axis=-1))


class Periodic(Kernel):
    """gpflow/kernels/periodic.py:28-111; uses the base kernel's active_dims."""

    def __init__(self, base_kernel: Stationary, period=1.0):
        super().__init__(None)
        self.base_kernel = base_kernel
        self.active_dims = base_kernel.active_dims
        self.period = period

    def K_diag(self, X):  # periodic.py:91-93
        return self.base_kernel.K_diag(X)

    def K(self, X, X2=None):  # periodic.py:95-111
        r = np.pi * difference_matrix(X, X2) / np.asarray(self.period, dtype=X.dtype)
        scaled_sine = np.sin(r) / np.asarray(self.base_kernel.lengthscales, dtype=X.dtype)
        if hasattr(self.base_kernel, "K_r"):
            return self.base_kernel.K_r(np.sum(np.abs(scaled_sine), -1))
        return self.base_kernel.K_r2(np.sum(np.square(scaled_sine), -1))


class ArcCosine(Kernel):
    """gpflow/kernels/misc.py:27-200."""

    def __init__(self, order=0, variance=1.0, weight_variances=1.0, bias_variance=1.0, active_dims=None):
        super().__init__(active_dims)
        if order not in (0, 1, 2):
            raise ValueError("Requested kernel order is not implemented.")
        self.order, self.variance, self.weight_variances, self.bias_variance = order, variance, weight_variances, bias_variance

    def _diag_weighted_product(self, X):  # misc.py:88-89
        return np.sum(np.asarray(self.weight_variances) * np.square(X), axis=-1) + self.bias_variance

    def _J(self, theta):  # misc.py:141-157
        if self.order == 0:
            return np.pi - theta
        if self.order == 1:
            return np.sin(theta) + (np.pi - theta) * np.cos(theta)
        return 3.0 * np.sin(theta) * np.cos(theta) + (np.pi - theta) * (1.0 + 2.0 * np.cos(theta) ** 2)

    def K(self, X, X2=None):  # misc.py:160-194
        Xd = np.sqrt(self._diag_weighted_product(X))
        X2_ = X if X2 is None else X2
        X2d = np.sqrt(self._diag_weighted_product(X2_))
        num = (np.asarray(self.weight_variances) * X) @ X2_.T + self.bias_variance
        cos_theta = num / Xd[:, None] / X2d[None, :]
        jitter = 1e-15
        theta = np.arccos(jitter + (1 - 2 * jitter) * cos_theta)
        return self.variance * (1.0 / np.pi) * self._J(theta) * Xd[:, None] ** self.order * X2d[None, :] ** self.order

    def K_diag(self, X):  # misc.py:197-200
        return self.variance * (1.0 / np.pi) * self._J(0.0) * self._diag_weighted_product(X) ** self.order


class Coregion(Kernel):
    """gpflow/kernels/misc.py:203-296."""

    def __init__(self, output_dim, rank, W=None, kappa=None, active_dims=None):
        super().__init__(active_dims)
        self.W = 0.1 * np.ones((output_dim, rank)) if W is None else np.asarray(W)
        self.kappa = np.ones(output_dim) if kappa is None else np.asarray(kappa)

    def output_covariance(self):
        return self.W @ self.W.T + np.diag(self.kappa)

    def K(self, X, X2=None):
        B = self.output_covariance()
        i = X[..., 0].astype(np.int32)
        j = i if X2 is None else X2[..., 0].astype(np.int32)
        return B[np.ix_(i, j)]

    def K_diag(self, X):
        return (np.sum(np.square(self.W), 1) + self.kappa)[X[..., 0].astype(np.int32)]


class ChangePoints(Kernel):
    """gpflow/kernels/changepoints.py:26-193 (1-D inputs)."""

    def __init__(self, kernels, locations, steepness=1.0):
        super().__init__(None)
        self.kernels, self.locations = list(kernels), np.asarray(locations, dtype=np.float64)
        self.steepness = steepness

    def _sigmoids(self, X):  # changepoints.py:189-193
        return 1.0 / (1.0 + np.exp(-np.asarray(self.steepness) * (X[:, :, None] - self.locations.reshape(1, 1, -1))))

    def __call__(self, X, X2=None, *, full_cov=True, presliced=False):
        return self.K(X, X2) if full_cov else self.K_diag(X)

    def K(self, X, X2=None):  # changepoints.py:86-149
        sig_X = self._sigmoids(X)[:, 0, :]                                      # [N, Ncp]
        sig_X2 = sig_X if X2 is None else self._sigmoids(X2)[:, 0, :]
        starters = sig_X[:, None, :] * sig_X2[None, :, :]
        stoppers = (1 - sig_X)[:, None, :] * (1 - sig_X2)[None, :, :]
        ones = np.ones(starters.shape[:2] + (1,), dtype=X.dtype)
        starters = np.concatenate([ones, starters], axis=-1)
        stoppers = np.concatenate([stoppers, ones], axis=-1)
        stack = np.stack([k(X, X2) for k in self.kernels], axis=-1)
        return np.sum(stack * starters * stoppers, axis=-1)

    def K_diag(self, X):  # changepoints.py:152-176
        sig = self._sigmoids(X)[:, 0, :]
        ones = np.ones((X.shape[0], 1), dtype=X.dtype)
        starters = np.concatenate([ones, sig * sig], axis=-1)
        stoppers = np.concatenate([(1 - sig) * (1 - sig), ones], axis=-1)
        stack = np.stack([k(X, full_cov=False) for k in self.kernels], axis=-1)
        return np.sum(stack * starters * stoppers, axis=-1)


class SeparateIndependent:
    """gpflow/kernels/multioutput/kernels.py:200-271, full_output_cov=False rows only."""

    def __init__(self, kernels: Sequence[Kernel]):
        self.kernels = list(kernels)

    def K(self, X, X2=None):  # kernels.py:236-239 -> [P, N, N2]
        return np.stack([k(X, X2) for k in self.kernels], axis=0)

    def K_diag(self, X):  # kernels.py:265-271 -> [N, P]
        return np.stack([k(X, full_cov=False) for k in self.kernels], axis=-1)


# ----------------------------------------------------------------------------------------
# covariances, noise
# ----------------------------------------------------------------------------------------
def Kuu(Z: np.ndarray, kernel: Kernel, *, jitter: float = 0.0) -> np.ndarray:
    """gpflow/covariances/kuus.py:24-34."""
    Kzz = kernel(Z)
    Kzz = Kzz + np.asarray(jitter, dtype=Kzz.dtype) * np.eye(Z.shape[0], dtype=Kzz.dtype)
    return Kzz


def Kuf(Z: np.ndarray, kernel: Kernel, Xnew: np.ndarray) -> np.ndarray:
    """gpflow/covariances/kufs.py:25-34 — [M, N], inducing first."""
    return kernel(Z, Xnew)


def add_noise_cov(K: np.ndarray, likelihood_variance) -> np.ndarray:
    """gpflow/utilities/model_utils.py:33-38 — diagonal shift, no jitter."""
    K = K.copy()
    idx = np.arange(K.shape[-1])
    K[..., idx, idx] = K[..., idx, idx] + np.asarray(likelihood_variance, dtype=K.dtype)
    return K


# ----------------------------------------------------------------------------------------
# mean functions (gpflow/functions.py:96-126,173-204)
# ----------------------------------------------------------------------------------------
class ZeroMean:
    def __init__(self, output_dim: int = 1):
        self.output_dim = output_dim

    def __call__(self, X):
        return np.zeros((X.shape[0], self.output_dim), dtype=X.dtype)


class ConstantMean:
    def __init__(self, c):
        self.c = np.atleast_1d(c)

    def __call__(self, X):
        return np.tile(self.c.astype(X.dtype).reshape(1, -1), (X.shape[0], 1))


class LinearMean:
    def __init__(self, A, b):
        self.A, self.b = np.atleast_2d(A), np.atleast_1d(b)

    def __call__(self, X):
        return X @ self.A.astype(X.dtype) + self.b.astype(X.dtype)


def _mean(mean_function, X, P):
    if mean_function is None:
        return np.zeros((X.shape[0], P), dtype=X.dtype)
    return mean_function(X)


# ----------------------------------------------------------------------------------------
# linear algebra helpers
# ----------------------------------------------------------------------------------------
def cholesky(A: np.ndarray) -> np.ndarray:
    """tf.linalg.cholesky call sites: lower factor, upper triangle zero."""
    return sla.cholesky(A, lower=True, check_finite=False)


def tri_solve(L: np.ndarray, B: np.ndarray, *, trans: bool = False) -> np.ndarray:
    """tf.linalg.triangular_solve(L, B, lower=True[, adjoint=True])."""
    return sla.solve_triangular(L, B, lower=True, trans=1 if trans else 0, check_finite=False)


def multivariate_normal(x: np.ndarray, mu: np.ndarray, L: np.ndarray) -> np.ndarray:
    """gpflow/logdensities.py:139-156 — one log-density per column."""
    d = x - mu
    alpha = tri_solve(L, d)
    num_dims = d.shape[0]
    p = -0.5 * np.sum(np.square(alpha), 0)
    p -= 0.5 * num_dims * np.asarray(LOG2PI, dtype=L.dtype)
    p -= np.sum(np.log(np.diag(L)))
    return p


# ----------------------------------------------------------------------------------------
# GPR
# ----------------------------------------------------------------------------------------
def gpr_log_marginal_likelihood(X, Y, kernel, noise_variance, mean_function=None) -> float:
    """gpflow/models/gpr.py:91-107."""
    K = kernel(X)
    ks = add_noise_cov(K, noise_variance)
    L = cholesky(ks)
    m = _mean(mean_function, X, Y.shape[1])
    log_prob = multivariate_normal(Y, m, L)
    return float(np.sum(log_prob))


def base_conditional_with_lm(Kmn, Lm, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    """gpflow/conditionals/util.py:84-169 (no leading batch dims)."""
    num_func = f.shape[-1]
    A = tri_solve(Lm, Kmn)  # util.py:125
    if full_cov:
        fvar = Knn - A.T @ A  # util.py:129
        fvar = np.broadcast_to(fvar[None], (num_func,) + fvar.shape).copy()
    else:
        fvar = Knn - np.sum(np.square(A), -2)  # util.py:133
        fvar = np.broadcast_to(fvar[None], (num_func,) + fvar.shape).copy()
    if not white:
        A = tri_solve(Lm, A, trans=True)  # util.py:138-139
    fmean = A.T @ f  # util.py:144
    if q_sqrt is not None:
        if q_sqrt.ndim == 2:
            LTA = A[None] * q_sqrt.T[:, :, None]  # util.py:149
        else:
            Lq = np.tril(q_sqrt)  # util.py:151 band_part(-1, 0)
            LTA = np.matmul(np.transpose(Lq, (0, 2, 1)), A)  # util.py:157  (L^T A per r)
        if full_cov:
            fvar = fvar + np.matmul(np.transpose(LTA, (0, 2, 1)), LTA)
        else:
            fvar = fvar + np.sum(np.square(LTA), -2)  # util.py:164
    if not full_cov:
        fvar = fvar.T  # util.py:167
    return fmean, fvar


def base_conditional(Kmn, Kmm, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    """gpflow/conditionals/util.py:37-70."""
    Lm = cholesky(Kmm)
    return base_conditional_with_lm(Kmn, Lm, Knn, f, full_cov=full_cov, q_sqrt=q_sqrt, white=white)


def gpr_predict_f(X, Y, kernel, noise_variance, Xnew, mean_function=None, full_cov=False):
    """gpflow/posteriors.py:384-443 (GPRPosterior) + posteriors.py:225-229 (mean add)."""
    err = Y - _mean(mean_function, X, Y.shape[1])
    Kmm = kernel(X)
    Lm = cholesky(add_noise_cov(Kmm, noise_variance))
    Knn = kernel(Xnew, full_cov=full_cov)
    Kmn = kernel(X, Xnew)
    mean, var = base_conditional_with_lm(Kmn, Lm, Knn, err, full_cov=full_cov, q_sqrt=None, white=False)
    return mean + _mean(mean_function, Xnew, Y.shape[1]), var


# ----------------------------------------------------------------------------------------
# SGPR
# ----------------------------------------------------------------------------------------
@dataclass
class SGPRCommon:
    sigma_sq: np.ndarray
    sigma: np.ndarray
    A: np.ndarray
    B: np.ndarray
    LB: np.ndarray
    AAT: np.ndarray
    L: np.ndarray


def sgpr_common(X, kernel, Z, noise_variance, jitter=DEFAULT_JITTER) -> SGPRCommon:
    """gpflow/models/sgpr.py:181-209."""
    N = X.shape[0]
    sigma_sq = np.broadcast_to(np.asarray(noise_variance, dtype=X.dtype), (N,)).copy()
    sigma = np.sqrt(sigma_sq)
    kuf = Kuf(Z, kernel, X)
    kuu = Kuu(Z, kernel, jitter=jitter)
    L = cholesky(kuu)
    A = tri_solve(L, kuf / sigma)
    AAT = A @ A.T
    B = add_noise_cov(AAT, 1.0)
    LB = cholesky(B)
    return SGPRCommon(sigma_sq, sigma, A, B, LB, AAT, L)


def sgpr_elbo(X, Y, kernel, Z, noise_variance, mean_function=None, jitter=DEFAULT_JITTER) -> float:
    """gpflow/models/sgpr.py:214-289."""
    c = sgpr_common(X, kernel, Z, noise_variance, jitter)
    N, P = Y.shape
    # logdet_term, sgpr.py:214-246
    kdiag = kernel(X, full_cov=False)
    trace_k = np.sum(kdiag / c.sigma_sq)
    trace_q = np.sum(np.diag(c.AAT))
    trace = trace_k - trace_q
    half_logdet_b = np.sum(np.log(np.diag(c.LB)))
    log_sigma_sq = np.sum(np.log(c.sigma_sq))
    logdet = -P * (half_logdet_b + 0.5 * log_sigma_sq + 0.5 * trace)
    # quad_term, sgpr.py:251-271
    err = (Y - _mean(mean_function, X, P)) / c.sigma[:, None]
    Aerr = c.A @ err
    cc = tri_solve(c.LB, Aerr)
    quad = -0.5 * (np.sum(np.square(err)) - np.sum(np.square(cc)))
    const = -0.5 * N * P * LOG2PI  # sgpr.py:286
    return float(const + logdet + quad)


def sgpr_upper_bound(X, Y, kernel, Z, noise_variance, mean_function=None, jitter=DEFAULT_JITTER) -> float:
    """gpflow/models/sgpr.py:87-147 (Titsias 2014 upper bound on the GPR marginal likelihood; the reference checks
    elbo < GPR lml < upper_bound in tests/integration/test_method_equivalence.py:297-327)."""
    N, P = Y.shape
    sigma_sq = np.broadcast_to(np.asarray(noise_variance, dtype=X.dtype), (N,)).copy()
    sigma = np.sqrt(sigma_sq)
    Kdiag = kernel(X, full_cov=False)
    kuu = Kuu(Z, kernel, jitter=jitter)
    kuf = Kuf(Z, kernel, X)
    I = np.eye(kuu.shape[0], dtype=X.dtype)
    L = cholesky(kuu)
    A = tri_solve(L, kuf)
    A_sigma = tri_solve(L, kuf / sigma)
    B = I + A_sigma @ A_sigma.T
    LB = cholesky(B)
    c = np.sum(Kdiag) - np.sum(np.square(A))        # trace bound, sgpr.py:126
    cn_var = sigma_sq + c
    cn_std = np.sqrt(cn_var)
    const = -0.5 * np.sum(np.log(2 * np.pi * sigma_sq))
    logdet = -np.sum(np.log(np.diag(LB)))
    A_cn = tri_solve(L, kuf / cn_std)
    err = Y - _mean(mean_function, X, P)
    LC = cholesky(I + A_cn @ A_cn.T)
    v = tri_solve(LC, A_cn @ (err / cn_std[:, None]))
    quad = -0.5 * np.sum(np.square(err / cn_std[:, None])) + 0.5 * np.sum(np.square(v))
    return float(const + logdet + quad)


def sgpr_predict_f(X, Y, kernel, Z, noise_variance, Xnew, mean_function=None, full_cov=False,
                   jitter=DEFAULT_JITTER):
    """gpflow/posteriors.py:479-551 (SGPRPosterior)."""
    P = Y.shape[1]
    c = sgpr_common(X, kernel, Z, noise_variance, jitter)
    err = Y - _mean(mean_function, X, P)
    Aerr = c.A @ (err / c.sigma[:, None])
    cc = tri_solve(c.LB, Aerr)
    Kus = Kuf(Z, kernel, Xnew)
    tmp1 = tri_solve(c.L, Kus)
    tmp2 = tri_solve(c.LB, tmp1)
    mean = tmp2.T @ cc
    if full_cov:
        var = kernel(Xnew) + tmp2.T @ tmp2 - tmp1.T @ tmp1
        var = np.tile(var[None], (P, 1, 1))
    else:
        var = kernel(Xnew, full_cov=False) + np.sum(np.square(tmp2), 0) - np.sum(np.square(tmp1), 0)
        var = np.tile(var[:, None], (1, P))
    return mean + _mean(mean_function, Xnew, P), var


def sgpr_compute_qu(X, Y, kernel, Z, noise_variance, mean_function=None, jitter=DEFAULT_JITTER):
    """gpflow/models/sgpr.py:346-377."""
    kuf = Kuf(Z, kernel, X)
    kuu = Kuu(Z, kernel, jitter=jitter)
    var = np.broadcast_to(np.asarray(noise_variance, dtype=X.dtype), (X.shape[0],))
    std = np.sqrt(var)
    scaled_kuf = kuf / std
    sig = kuu + scaled_kuf @ scaled_kuf.T
    sig_sqrt = cholesky(sig)
    sig_sqrt_kuu = tri_solve(sig_sqrt, kuu)
    cov = sig_sqrt_kuu.T @ sig_sqrt_kuu
    err = Y - _mean(mean_function, X, Y.shape[1])
    scaled_err = err / std[:, None]
    mu = sig_sqrt_kuu.T @ tri_solve(sig_sqrt, scaled_kuf @ scaled_err)
    return mu, cov


# ----------------------------------------------------------------------------------------
# KL, Gaussian likelihood, SVGP
# ----------------------------------------------------------------------------------------
def gauss_kl(q_mu, q_sqrt, K=None, *, K_cholesky=None) -> float:
    """gpflow/kullback_leiblers.py:59-165."""
    if (K is not None) and (K_cholesky is not None):
        raise ValueError("Ambiguous arguments: gauss_kl() must only be passed one of `K` or `K_cholesky`.")
    is_white = (K is None) and (K_cholesky is None)
    is_diag = q_sqrt.ndim == 2
    M, L = q_mu.shape
    is_batched = False
    if is_white:
        alpha = q_mu
    else:
        Lp = np.stack([cholesky(k) for k in K]) if (K is not None and K.ndim == 3) else (
            cholesky(K) if K is not None else K_cholesky)
        is_batched = Lp.ndim == 3
        if is_batched:
            alpha = np.stack([tri_solve(Lp[l], q_mu[:, l : l + 1]) for l in range(L)])
        else:
            alpha = tri_solve(Lp, q_mu)
    if is_diag:
        Lq = Lq_diag = q_sqrt
        Lq_full = np.stack([np.diag(q_sqrt[:, l]) for l in range(L)])
    else:
        Lq = Lq_full = np.tril(q_sqrt)
        Lq_diag = np.stack([np.diag(Lq[l]) for l in range(L)], axis=1)
    mahalanobis = np.sum(np.square(alpha))  # :124
    constant = -float(q_mu.size)  # :127
    logdet_qcov = np.sum(np.log(np.square(Lq_diag)))  # :130
    if is_white:
        trace = np.sum(np.square(Lq))  # :134
    else:
        if is_diag and not is_batched:  # :136-145
            Lp_inv = tri_solve(Lp, np.eye(M, dtype=Lp.dtype))
            K_inv = np.diag(sla.solve_triangular(Lp.T, Lp_inv, lower=False))[:, None]
            trace = np.sum(K_inv * np.square(q_sqrt))
        else:  # :152-153
            if is_batched:
                LpiLq = np.stack([tri_solve(Lp[l], Lq_full[l]) for l in range(L)])
            else:
                LpiLq = np.stack([tri_solve(Lp, Lq_full[l]) for l in range(L)])
            trace = np.sum(np.square(LpiLq))
    twoKL = mahalanobis + constant - logdet_qcov + trace
    if not is_white:  # :158-163
        if is_batched:
            s = sum(np.sum(np.log(np.square(np.diag(Lp[l])))) for l in range(L))
            twoKL += s
        else:
            twoKL += L * np.sum(np.log(np.square(np.diag(Lp))))
    return float(0.5 * twoKL)


def prior_kl(Z, kernel, q_mu, q_sqrt, whiten=False, jitter=DEFAULT_JITTER) -> float:
    """gpflow/kullback_leiblers.py:31-49."""
    if whiten:
        return gauss_kl(q_mu, q_sqrt, None)
    return gauss_kl(q_mu, q_sqrt, Kuu(Z, kernel, jitter=jitter))


def gaussian_variational_expectations(Fmu, Fvar, Y, variance):
    """gpflow/likelihoods/scalar_continuous.py:139-148 -> [N]."""
    variance = np.asarray(variance, dtype=Fmu.dtype)
    return np.sum(-0.5 * LOG2PI - 0.5 * np.log(variance) - 0.5 * ((Y - Fmu) ** 2 + Fvar) / variance, axis=-1)


def gaussian_predict_mean_and_var(Fmu, Fvar, variance):
    """gpflow/likelihoods/scalar_continuous.py:127-130."""
    return Fmu, Fvar + np.asarray(variance, dtype=Fmu.dtype)


def gaussian_predict_log_density(Fmu, Fvar, Y, variance):
    """gpflow/likelihoods/scalar_continuous.py:133-136 with logdensities.py:29-30."""
    var = Fvar + np.asarray(variance, dtype=Fmu.dtype)
    return np.sum(-0.5 * (LOG2PI + np.log(var) + np.square(Fmu - Y) / var), axis=-1)


def svgp_predict_f(Xnew, Z, kernel, q_mu, q_sqrt, *, whiten=True, full_cov=False,
                   mean_function=None, jitter=DEFAULT_JITTER):
    """gpflow/posteriors.py:827-841 (IndependentPosteriorSingleOutput._conditional_fused)."""
    Knn = kernel(Xnew, full_cov=full_cov)
    Kmm = Kuu(Z, kernel, jitter=jitter)
    Kmn = Kuf(Z, kernel, Xnew)
    fmean, fvar = base_conditional(Kmn, Kmm, Knn, q_mu, full_cov=full_cov, q_sqrt=q_sqrt, white=whiten)
    return fmean + _mean(mean_function, Xnew, q_mu.shape[1]), fvar


def svgp_elbo(Xb, Yb, Z, kernel, q_mu, q_sqrt, noise_variance, *, whiten=True, num_data=None,
              mean_function=None, jitter=DEFAULT_JITTER) -> float:
    """gpflow/models/svgp.py:166-181."""
    kl = prior_kl(Z, kernel, q_mu, q_sqrt, whiten=whiten, jitter=jitter)
    f_mean, f_var = svgp_predict_f(Xb, Z, kernel, q_mu, q_sqrt, whiten=whiten, full_cov=False,
                                   mean_function=mean_function, jitter=jitter)
    var_exp = gaussian_variational_expectations(f_mean, f_var, Yb, noise_variance)
    scale = 1.0 if num_data is None else float(num_data) / Xb.shape[0]
    return float(np.sum(var_exp) * scale - kl)


def svgp_cached_alpha_qinv(Z, kernel, q_mu, q_sqrt, *, whiten=True, jitter=DEFAULT_JITTER):
    """gpflow/posteriors.py:694-746 (BasePosterior._precompute), single-output kernel."""
    kuu = Kuu(Z, kernel, jitter=jitter)
    L = cholesky(kuu)
    M, P = q_mu.shape
    if not whiten:
        alpha = tri_solve(L, tri_solve(L, q_mu), trans=True)
    else:
        alpha = tri_solve(L, q_mu, trans=True)
    I = np.eye(M, dtype=L.dtype)
    if q_sqrt is None:
        B = np.broadcast_to(I, (P, M, M))
    else:
        qs = np.stack([np.diag(q_sqrt[:, p]) for p in range(P)]) if q_sqrt.ndim == 2 else q_sqrt
        if not whiten:
            Linv_q = np.stack([tri_solve(L, qs[p]) for p in range(P)])
            C = np.einsum("pij,pkj->pik", Linv_q, Linv_q)
        else:
            C = np.einsum("pij,pkj->pik", qs, qs)
        B = I[None] - C
    Qinv = np.stack([tri_solve(L, tri_solve(L, B[p], trans=True).T, trans=True) for p in range(P)])
    return alpha, Qinv


def svgp_predict_f_cached(Xnew, Z, kernel, alpha, Qinv, mean_function=None):
    """gpflow/posteriors.py:794-822 (full_cov=False)."""
    kuf = Kuf(Z, kernel, Xnew)
    Kff = kernel(Xnew, full_cov=False)
    mean = kuf.T @ alpha
    cov = Kff[None] - np.sum(kuf[None] * (Qinv @ kuf), axis=-2)
    return mean + _mean(mean_function, Xnew, alpha.shape[1]), cov.T


# ----------------------------------------------------------------------------------------
# sampling, FITC, multi-output posteriors
# ----------------------------------------------------------------------------------------
def sample_mvn(mean, cov, full_cov: bool, eps, jitter=DEFAULT_JITTER):
    """gpflow/conditionals/util.py:179-211 with the standard-normal draws `eps` given ([S, N, D] when not full_cov,
    [N, D, S] when full_cov) instead of tf.random.normal; returns [S, N, D]."""
    if not full_cov:
        return mean[None] + np.sqrt(cov)[None] * eps
    D = mean.shape[-1]
    chol = np.stack([cholesky(c + jitter * np.eye(D)) for c in cov])          # [N, D, D]
    samples = mean[..., None] + chol @ eps                                    # [N, D, S]
    return np.transpose(samples, (2, 0, 1))


def predict_f_samples(mean, cov, full_cov: bool, eps, full_output_cov=False, jitter=DEFAULT_JITTER):
    """gpflow/models/model.py:267-288 given predict_f's (mean [N, P], cov); [S, N, P]."""
    if full_cov:
        s = sample_mvn(mean.T, cov, True, eps, jitter)                        # [S, P, N]
        return np.transpose(s, (0, 2, 1))
    return sample_mvn(mean, cov, full_output_cov, eps, jitter)


def gprfitc_common(X, Y, kernel, Z, noise_variance, mean_function=None, jitter=DEFAULT_JITTER):
    """gpflow/models/sgpr.py:399-432."""
    err = Y - _mean(mean_function, X, Y.shape[1])
    Kdiag = kernel(X, full_cov=False)
    kuf = Kuf(Z, kernel, X)
    kuu = Kuu(Z, kernel, jitter=jitter)
    sigma_sq = np.broadcast_to(np.asarray(noise_variance, dtype=X.dtype), (X.shape[0],))
    Luu = cholesky(kuu)
    V = tri_solve(Luu, kuf)
    nu = Kdiag - np.sum(np.square(V), 0) + sigma_sq
    B = np.eye(Z.shape[0], dtype=X.dtype) + (V / nu) @ V.T
    L = cholesky(B)
    beta = err / nu[:, None]
    alpha = V @ beta
    gamma = tri_solve(L, alpha)
    return err, nu, Luu, L, alpha, beta, gamma


def gprfitc_lml(X, Y, kernel, Z, noise_variance, mean_function=None, jitter=DEFAULT_JITTER) -> float:
    """gpflow/models/sgpr.py:440-480."""
    err, nu, _, L, _, _, gamma = gprfitc_common(X, Y, kernel, Z, noise_variance, mean_function, jitter)
    maha = -0.5 * np.sum(np.square(err) / nu[:, None]) + 0.5 * np.sum(np.square(gamma))
    const = -0.5 * X.shape[0] * np.log(2.0 * np.pi)
    logdet = -0.5 * np.sum(np.log(nu)) - np.sum(np.log(np.diag(L)))
    return float(maha + (const + logdet) * Y.shape[1])


def gprfitc_predict_f(X, Y, kernel, Z, noise_variance, Xnew, mean_function=None, full_cov=False, jitter=DEFAULT_JITTER):
    """gpflow/models/sgpr.py:482-523."""
    _, _, Luu, L, _, _, gamma = gprfitc_common(X, Y, kernel, Z, noise_variance, mean_function, jitter)
    Kus = Kuf(Z, kernel, Xnew)
    w = tri_solve(Luu, Kus)
    tmp = tri_solve(L, gamma, trans=True)
    mean = w.T @ tmp + _mean(mean_function, Xnew, Y.shape[1])
    iA = tri_solve(L, w)
    P = Y.shape[1]
    if full_cov:
        var = kernel(Xnew) - w.T @ w + iA.T @ iA
        return mean, np.tile(var[None], (P, 1, 1))
    var = kernel(Xnew, full_cov=False) - np.sum(np.square(w), 0) + np.sum(np.square(iA), 0)
    return mean, np.tile(var[:, None], (1, P))


def mo_independent_predict_f(Xnew, Zs, kernels, q_mu, q_sqrt, *, whiten=True, full_cov=False, jitter=DEFAULT_JITTER):
    """gpflow/posteriors.py:844-885 / conditionals/util.py:566-629: L independent latent GPs, latent l with inducing
    inputs Zs[l] and kernel kernels[l] (lists of length 1 are shared).  fmean [N, L]; fvar [N, L] or [L, N, N]."""
    L = q_mu.shape[1]
    means, vars_ = [], []
    for l in range(L):
        Z = Zs[l if len(Zs) > 1 else 0]
        k = kernels[l if len(kernels) > 1 else 0]
        qs = None if q_sqrt is None else (q_sqrt[:, l:l + 1] if q_sqrt.ndim == 2 else q_sqrt[l:l + 1])
        m, v = base_conditional(Kuf(Z, k, Xnew), Kuu(Z, k, jitter=jitter), k(Xnew, full_cov=full_cov), q_mu[:, l:l + 1],
                                full_cov=full_cov, q_sqrt=qs, white=whiten)
        means.append(m[:, 0])
        vars_.append(v[0] if full_cov else v[:, 0])
    return np.stack(means, 1), (np.stack(vars_, 0) if full_cov else np.stack(vars_, 1))


def mix_latent_gp(W, g_mean, g_var, full_cov: bool, full_output_cov: bool):
    """gpflow/conditionals/util.py:518-563; W [P, L]."""
    f_mean = g_mean @ W.T
    if full_cov and full_output_cov:        # g_var [L, N, N] -> [N, P, N, P]
        return f_mean, np.einsum("lnm,pl,ql->npmq", g_var, W, W)
    if full_cov:                            # -> [P, N, N]
        return f_mean, np.einsum("lnm,pl->pnm", g_var, W ** 2)
    if full_output_cov:                     # g_var [N, L] -> [N, P, P]
        return f_mean, np.einsum("nl,pl,ql->npq", g_var, W, W)
    return f_mean, g_var @ (W ** 2).T


# ----------------------------------------------------------------------------------------
# multi-output GPR sum (config 5: SeparateIndependent semantics, one GPR per output)
# ----------------------------------------------------------------------------------------
def separate_gpr_lml(X, Y, kernels: Sequence[Kernel], noise_variance) -> float:
    """Sum_p GPR_p.lml with k_p per output; gpflow/kernels/multioutput/kernels.py:236-239 stacks
    independent per-output problems and gpr.py:105-107 sums per-column log-probs."""
    return float(sum(gpr_log_marginal_likelihood(X, Y[:, p : p + 1], kernels[p], noise_variance)
                     for p in range(len(kernels))))


# ----------------------------------------------------------------------------------------
# deterministic synthetic inputs; shared by oracle, tests and bench
# ----------------------------------------------------------------------------------------
def make_data(config_index: int, N: int, D: int, P: int, M: int = 0, n_new: int = 0, dtype=np.float64):
    rng = np.random.default_rng(20220523 + config_index)
    X = rng.standard_normal((N, D))
    w = rng.standard_normal((D, P)) / np.sqrt(D)
    Y = np.sin(X @ w) + 0.1 * rng.standard_normal((N, P))
    out = {"X": X.astype(dtype), "Y": Y.astype(dtype)}
    if M:
        out["Z"] = X[rng.choice(N, M, replace=False)].astype(dtype)
    if n_new:
        out["Xnew"] = rng.standard_normal((n_new, D)).astype(dtype)
    return out


def make_q(config_index: int, M: int, P: int, dtype=np.float64):
    rng = np.random.default_rng(7 + 20220523 + config_index)
    q_mu = 0.1 * rng.standard_normal((M, P))
    q_sqrt = np.stack([np.tril(0.1 * rng.standard_normal((M, M))) + np.eye(M) for _ in range(P)])
    return q_mu.astype(dtype), q_sqrt.astype(dtype)


# ----------------------------------------------------------------------------------------
# VGP (sibling model on the same operators)
# ----------------------------------------------------------------------------------------
def vgp_elbo(X, Y, kernel, q_mu, q_sqrt, noise_variance, mean_function=None, jitter=DEFAULT_JITTER) -> float:
    """gpflow/models/vgp.py:111-143 with a Gaussian likelihood: whitened q(v) = N(q_mu, q_sqrt q_sqrt^T), f = L v + m."""
    N, P = Y.shape
    KL = gauss_kl(q_mu, q_sqrt)
    K = kernel(X) + np.eye(N, dtype=X.dtype) * np.asarray(jitter, dtype=X.dtype)
    L = cholesky(K)
    fmean = L @ q_mu + _mean(mean_function, X, P)
    q_sqrt_dnn = np.tril(q_sqrt)                       # band_part(q_sqrt, -1, 0)  [P, N, N]
    LTA = np.matmul(L[None], q_sqrt_dnn)               # [P, N, N]
    fvar = np.sum(np.square(LTA), axis=2).T            # [N, P]
    var_exp = gaussian_variational_expectations(fmean, fvar, Y, noise_variance)
    return float(np.sum(var_exp) - KL)


def vgp_predict_f(X, kernel, q_mu, q_sqrt, Xnew, mean_function=None, full_cov=False, jitter=DEFAULT_JITTER):
    """gpflow/models/vgp.py:145-161 -> conditionals/conditionals.py:39-116 (Kmm = K(X) + jitter I, white=True)."""
    P = q_mu.shape[1]
    Kmm = kernel(X) + np.eye(X.shape[0], dtype=X.dtype) * np.asarray(jitter, dtype=X.dtype)
    Kmn = kernel(X, Xnew)
    Knn = kernel(Xnew, full_cov=full_cov)
    mean, var = base_conditional(Kmn, Kmm, Knn, q_mu, full_cov=full_cov, q_sqrt=q_sqrt, white=True)
    return mean + _mean(mean_function, Xnew, P), var
