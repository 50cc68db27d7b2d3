"""Kuu / Kuf dispatchers (mirrors gpflow/covariances/dispatch.py:17-18, kuus.py:24-34, kufs.py:25-34)."""
from __future__ import annotations

from . import _lib, ops
from .inducing_variables import InducingPoints
from .kernels import Kernel, compile_kernel
from .utilities.multipledispatch import Dispatcher

Kuu = Dispatcher("Kuu")
Kuf = Dispatcher("Kuf")


@Kuu.register(InducingPoints, Kernel)
def _Kuu_points(inducing_variable: InducingPoints, kernel: Kernel, *, jitter: float = 0.0):
    """kernel(Z) + jitter*I, built in one pass with the jitter fused on the diagonal (kuus.py:29-34)."""
    Z = ops.to_device(inducing_variable.Z)
    return ops.kbuild(compile_kernel(kernel, Z.shape[1]), Z, None, diag_scalar=jitter)


@Kuf.register(InducingPoints, Kernel, object)
def _Kuf_points(inducing_variable: InducingPoints, kernel: Kernel, Xnew):
    """kernel(Z, Xnew) -> [M, N], inducing first (kufs.py:31-34)."""
    Z = ops.to_device(inducing_variable.Z)
    X = ops.to_device(Xnew)
    return ops.kbuild(compile_kernel(kernel, Z.shape[1]), Z, X)
