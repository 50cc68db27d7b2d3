"""Kuu / Kuf dispatchers (mirrors gpflow/covariances/dispatch.py:17-18, kuus.py:24-34, kufs.py:25-34)."""
from __future__ import annotations

from . import _lib, ops
from .inducing_variables import (FallbackSeparateIndependentInducingVariables, FallbackSharedIndependentInducingVariables,
                                 InducingPoints)
from .kernels import Kernel, MultioutputKernel, SeparateIndependent, SharedIndependent, kernel_matrix
from .utilities.multipledispatch import Dispatcher

Kuu = Dispatcher("Kuu")
Kuf = Dispatcher("Kuf")


@Kuu.register(InducingPoints, Kernel)
def _Kuu_points(inducing_variable: InducingPoints, kernel: Kernel, *, jitter: float = 0.0):
    """kernel(Z) + jitter*I, built in one pass with the jitter fused on the diagonal (kuus.py:29-34)."""
    Z = ops.to_device(inducing_variable.Z)
    return kernel_matrix(kernel, Z, None, diag_scalar=jitter)


@Kuf.register(InducingPoints, Kernel, object)
def _Kuf_points(inducing_variable: InducingPoints, kernel: Kernel, Xnew):
    """kernel(Z, Xnew) -> [M, N], inducing first (kufs.py:31-34)."""
    Z = ops.to_device(inducing_variable.Z)
    X = ops.to_device(Xnew)
    return kernel_matrix(kernel, Z, X)


# ---- multi-output (gpflow/covariances/multioutput/kuus.py:33-122, kufs.py:33-145): [M, M] / [M, N] when both the kernel and
# the inducing variables are shared, else one matrix per latent GP, [L, M, M] / [L, M, N] ---------------------------------
def _latent_pairs(inducing_variable, kernel: MultioutputKernel):
    """(inducing variables, kernel) of every latent GP; a single kernel / a single set of inducing variables is shared."""
    ks = kernel.latent_kernels
    ivs = inducing_variable.inducing_variables
    L = len(ivs) if len(ivs) > 1 else (len(ks) if len(ks) > 1 else kernel.num_latent_gps)
    if len(ks) not in (1, L) or len(ivs) not in (1, L):
        raise ValueError(f"{len(ks)} latent kernels do not match {len(ivs)} sets of inducing variables")
    return [(ivs[l if len(ivs) > 1 else 0], ks[l if len(ks) > 1 else 0]) for l in range(L)]


@Kuu.register(FallbackSharedIndependentInducingVariables, SharedIndependent)
def _Kuu_shared_shared(inducing_variable, kernel: SharedIndependent, *, jitter: float = 0.0):
    return Kuu(inducing_variable.inducing_variable, kernel.kernel, jitter=jitter)          # kuus.py:33-41 -> [M, M]


@Kuf.register(FallbackSharedIndependentInducingVariables, SharedIndependent, object)
def _Kuf_shared_shared(inducing_variable, kernel: SharedIndependent, Xnew):
    return Kuf(inducing_variable.inducing_variable, kernel.kernel, Xnew)                   # kufs.py:33-41 -> [M, N]


def _stack(mats):
    out = ops.empty((len(mats),) + tuple(mats[0].shape), like=mats[0])
    for l, m in enumerate(mats):
        ops.axpby(1.0, m, 0.0, out[l])
    return out


@Kuu.register((FallbackSharedIndependentInducingVariables, FallbackSeparateIndependentInducingVariables), MultioutputKernel)
def _Kuu_per_latent(inducing_variable, kernel: MultioutputKernel, *, jitter: float = 0.0):
    if isinstance(inducing_variable, FallbackSharedIndependentInducingVariables) and isinstance(kernel, SharedIndependent):
        return _Kuu_shared_shared(inducing_variable, kernel, jitter=jitter)
    return _stack([Kuu(iv, k, jitter=jitter) for iv, k in _latent_pairs(inducing_variable, kernel)])  # [L, M, M]


@Kuf.register((FallbackSharedIndependentInducingVariables, FallbackSeparateIndependentInducingVariables), MultioutputKernel,
              object)
def _Kuf_per_latent(inducing_variable, kernel: MultioutputKernel, Xnew):
    if isinstance(inducing_variable, FallbackSharedIndependentInducingVariables) and isinstance(kernel, SharedIndependent):
        return _Kuf_shared_shared(inducing_variable, kernel, Xnew)
    return _stack([Kuf(iv, k, Xnew) for iv, k in _latent_pairs(inducing_variable, kernel)])          # [L, M, N]
