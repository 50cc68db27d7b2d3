"""Kernels that are not functions of a Gram term: Cosine, Periodic, ArcCosine, Coregion and
ChangePoints (mirrors gpflow/kernels/stationaries.py:316-332, periodic.py:28-111, misc.py:27-296, changepoints.py:26-193).

They have no `gpk_knode` record for the fused K-build; each is evaluated by its own device kernel (`gpk_kaux`,
csrc/kaux.cu) and combines with other kernels through elementwise device ops (`Combination._materialise`).  Same classes,
constructor arguments and error behaviour as the reference."""
from __future__ import annotations

from typing import Any, Optional, Sequence

import numpy as np

from .. import _lib, ops
from ..base import Parameter, positive
from .base import ActiveDims, Combination, Kernel
from .stationaries import IsotropicStationary, Stationary


class _Materialised(Kernel):
    def is_fusable(self) -> bool:
        return False

    def _desc(self, D: int) -> _lib.KAux:
        raise NotImplementedError

    def _dims(self, D: int) -> np.ndarray:
        d = self._resolved_dims(D)
        d = np.arange(D) if d is None else d
        if len(d) > _lib.GPK_KAUX_MAXD:
            raise ValueError(f"{type(self).__name__}: at most {_lib.GPK_KAUX_MAXD} active dims")
        if np.any(d < 0) or np.any(d >= D):
            raise ValueError(f"active_dims {d} out of range for inputs with {D} columns")
        return d

    @staticmethod
    def _fill(desc: _lib.KAux, dims, scale=None, period=None) -> None:
        desc.n_dims = len(dims)
        for i, c in enumerate(dims):
            desc.dims[i] = int(c)
            desc.scale[i] = float(scale[i]) if scale is not None else 1.0
            desc.period[i] = float(period[i]) if period is not None else 1.0

    def _materialise(self, X, X2, full_cov: bool):
        lib = _lib.load()
        desc = self._desc(X.shape[-1])
        dc = ops.dtype_code(X)
        N = X.shape[0]
        if not full_cov:
            out = ops.empty((N,), like=X)
            _lib.check(lib.gpk_kaux_diag(desc, ops._p(X), N, ops._ld(X), ops._p(out), dc, ops._stream()), "gpk_kaux_diag")
            return out
        N2 = N if X2 is None else X2.shape[0]
        K = ops.empty((N, N2), like=X)
        _lib.check(lib.gpk_kaux(desc, ops._p(X), N, ops._ld(X), ops._p(X2), N2, ops._ld(X2) if X2 is not None else 0,
                                ops._p(K), ops._ld(K), dc, ops._stream()), "gpk_kaux")
        return K


def _per_dim(value: np.ndarray, n: int) -> np.ndarray:
    v = np.asarray(value, dtype=np.float64)
    return np.full(n, float(v)) if v.ndim == 0 else v.reshape(-1)


class AnisotropicStationary(Stationary, _Materialised):
    """stationaries.py:133-196: kernels of the per-dimension scaled differences (X - X2) / lengthscales."""

    def is_fusable(self) -> bool:
        return False


class Cosine(AnisotropicStationary):
    """k = sigma^2 cos(2 pi sum_d (x_d - x'_d) / l_d)  (stationaries.py:316-332)."""

    def _desc(self, D: int) -> _lib.KAux:
        dims = self._dims(D)
        ls = _per_dim(self.lengthscales.numpy(), len(dims))
        if len(ls) != len(dims):
            raise ValueError(f"Size of ARD parameter ({len(ls)}) does not match active dims ({len(dims)})")
        d = _lib.KAux(op=_lib.KAUX_COSINE, variance=float(self.variance.numpy()))
        self._fill(d, dims, scale=1.0 / ls)
        return d


class Periodic(_Materialised):
    """periodic.py:28-111: wraps an IsotropicStationary base kernel; uses the base kernel's active_dims."""

    def __init__(self, base_kernel: IsotropicStationary, period: Any = 1.0) -> None:
        if not isinstance(base_kernel, IsotropicStationary):
            raise TypeError("Periodic requires an IsotropicStationary kernel as the `base_kernel`")  # periodic.py:66-67
        super().__init__()
        self.base_kernel = base_kernel
        self.period = Parameter(period, transform=positive())
        self.base_kernel._validate_ard_active_dims(self.period)

    @property
    def active_dims(self):
        return self.base_kernel.active_dims

    @active_dims.setter
    def active_dims(self, value: ActiveDims) -> None:
        if hasattr(self, "base_kernel"):
            self.base_kernel.active_dims = value

    def _desc(self, D: int) -> _lib.KAux:
        b = self.base_kernel
        d0 = b._resolved_dims(D)
        dims = np.arange(D) if d0 is None else d0
        if len(dims) > _lib.GPK_KAUX_MAXD:
            raise ValueError(f"Periodic: at most {_lib.GPK_KAUX_MAXD} active dims")
        ls = _per_dim(b.lengthscales.numpy(), len(dims))
        per = _per_dim(self.period.numpy(), len(dims))
        if len(ls) != len(dims) or len(per) != len(dims):
            raise ValueError("Size of ARD parameter does not match active dims")
        d = _lib.KAux(op=_lib.KAUX_PERIODIC, base=b._op, variance=float(b.variance.numpy()),
                      alpha=float(getattr(b, "alpha", Parameter(1.0)).numpy()))
        self._fill(d, dims, scale=1.0 / ls, period=per)
        return d


class ArcCosine(_Materialised):
    """misc.py:27-200 (Cho & Saul 2009), orders 0, 1, 2."""

    implemented_orders = {0, 1, 2}

    def __init__(self, order: int = 0, variance: Any = 1.0, weight_variances: Any = 1.0, bias_variance: Any = 1.0, *,
                 active_dims: ActiveDims = None, name: Optional[str] = None) -> None:
        super().__init__(active_dims=active_dims, name=name)
        if order not in self.implemented_orders:
            raise ValueError("Requested kernel order is not implemented.")  # misc.py:67-68
        self.order = order
        self.variance = Parameter(variance, transform=positive())
        self.bias_variance = Parameter(bias_variance, transform=positive())
        self.weight_variances = Parameter(weight_variances, transform=positive())
        self._validate_ard_active_dims(self.weight_variances)

    @property
    def ard(self) -> bool:
        return self.weight_variances.numpy().ndim > 0

    def _desc(self, D: int) -> _lib.KAux:
        dims = self._dims(D)
        w = _per_dim(self.weight_variances.numpy(), len(dims))
        if len(w) != len(dims):
            raise ValueError(f"Size of ARD parameter ({len(w)}) does not match active dims ({len(dims)})")
        d = _lib.KAux(op=_lib.KAUX_ARCCOS, order=int(self.order), variance=float(self.variance.numpy()),
                      bias=float(self.bias_variance.numpy()))
        self._fill(d, dims, scale=w)
        return d


class Coregion(_Materialised):
    """misc.py:203-296: K(x, y) = B[x, y], B = W W^T + diag(kappa); the input column holds integer output indices."""

    def __init__(self, output_dim: int, rank: int, *, active_dims: ActiveDims = None, name: Optional[str] = None) -> None:
        super().__init__(active_dims=active_dims, name=name)
        self.output_dim = output_dim
        self.rank = rank
        self.W = Parameter(0.1 * np.ones((output_dim, rank)))
        self.kappa = Parameter(np.ones(output_dim), transform=positive())

    def output_covariance(self) -> np.ndarray:  # misc.py:249-251
        W = self.W.numpy()
        return W @ W.T + np.diag(self.kappa.numpy())

    def output_variance(self) -> np.ndarray:  # misc.py:256-258
        return np.sum(np.square(self.W.numpy()), 1) + self.kappa.numpy()

    def _desc(self, D: int) -> _lib.KAux:
        dims = self._dims(D)
        if len(dims) != 1:
            raise ValueError("The `Coregion` kernel requires a 1D input space.")  # misc.py:262
        B = np.ascontiguousarray(self.output_covariance(), dtype=np.float64)
        self._table = ops.torch().from_numpy(B).to(ops.require_cuda())  # kept alive for the launch
        d = _lib.KAux(op=_lib.KAUX_COREGION, table_dim=int(self.output_dim), table=self._table.data_ptr())
        self._fill(d, dims)
        return d


class ChangePoints(Combination):
    """changepoints.py:26-193: K = sum_i diag(a_i(X)) K_i(X, X2) diag(a_i(X2)) with a_i = sig_i (1 - sig_{i+1})
    (sig_0 = 1, 1 - sig_{Ncp+1} = 1), sig_i(x) = 1 / (1 + exp(-s_i (x - x0_i)))."""

    def __init__(self, kernels: Sequence[Kernel], locations: Any, steepness: Any = 1.0, name: Optional[str] = None):
        if len(kernels) != len(locations) + 1:
            raise ValueError("Number of kernels ({nk}) must be one more than the number of "
                             "changepoint locations ({nl})".format(nk=len(kernels), nl=len(locations)))
        if isinstance(steepness, Sequence) and len(steepness) != len(locations):
            raise ValueError("Dimension of steepness ({ns}) does not match number of changepoint "
                             "locations ({nl})".format(ns=len(steepness), nl=len(locations)))
        super().__init__(kernels, name=name)
        self.kernels = list(kernels)  # _set_kernels: no flattening (changepoints.py:82-84)
        self.locations = Parameter(locations)
        self.steepness = Parameter(steepness, transform=positive())

    def is_fusable(self) -> bool:
        return False

    def _weights(self, X, i: int):
        lib = _lib.load()
        loc = self.locations.numpy().reshape(-1)
        st = _per_dim(self.steepness.numpy(), len(loc))
        ncp = len(loc)
        out = ops.empty((X.shape[0],), like=X)
        has_lo, has_hi = int(i > 0), int(i < ncp)
        _lib.check(lib.gpk_changepoint_weights(ops._p(X), X.shape[0], ops._ld(X), 0, has_lo,
                                               float(loc[i - 1]) if has_lo else 0.0, float(st[i - 1]) if has_lo else 1.0,
                                               has_hi, float(loc[i]) if has_hi else 0.0, float(st[i]) if has_hi else 1.0,
                                               ops._p(out), ops.dtype_code(X), ops._stream()), "gpk_changepoint_weights")
        return out

    def _materialise(self, X, X2, full_cov: bool):
        if X.shape[-1] != 1:
            raise ValueError("The `ChangePoints` kernel requires a 1D input space.")  # changepoints.py:88
        acc = None
        for i, k in enumerate(self.kernels):
            a = self._weights(X, i)
            if full_cov:
                Ki = k(X, X2)
                ops.scale_rows_(Ki, a)
                ops.scale_cols_(Ki, a if X2 is None else self._weights(X2, i))
            else:
                Ki = k(X, full_cov=False).view(-1, 1)
                ops.scale_rows_(Ki, a)
                ops.scale_rows_(Ki, a)
                Ki = Ki.view(-1)
            if acc is None:
                acc = Ki  # a fresh tensor: k(...) allocates its result
            else:
                ops.axpby(1.0, Ki, 1.0, acc)
        return acc
