"""Independent multi-output kernels, full_output_cov=False rows only
(mirrors gpflow/kernels/multioutput/kernels.py:118-271)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

from .. import ops
from .base import Combination, Kernel, compile_kernel


class MultioutputKernel(Kernel):
    @property
    def num_latent_gps(self) -> int:
        raise NotImplementedError

    @property
    def latent_kernels(self) -> Tuple[Kernel, ...]:
        raise NotImplementedError


class SharedIndependent(MultioutputKernel):
    """One kernel shared by `output_dim` independent outputs (kernels.py:118-197)."""

    def __init__(self, kernel: Kernel, output_dim: int) -> None:
        super().__init__()
        self.kernel = kernel
        self.output_dim = output_dim

    @property
    def num_latent_gps(self) -> int:
        return self.output_dim

    @property
    def latent_kernels(self) -> Tuple[Kernel, ...]:
        return (self.kernel,)

    def K(self, X, X2=None, full_output_cov: bool = False):
        if full_output_cov:
            raise NotImplementedError("full_output_cov=True is outside the hot path")
        return self.kernel(X, X2)  # [N, N2], broadcast over outputs by the caller

    def K_diag(self, X, full_output_cov: bool = False):
        if full_output_cov:
            raise NotImplementedError("full_output_cov=True is outside the hot path")
        return self.kernel(X, full_cov=False)

    def __call__(self, X, X2=None, *, full_cov: bool = True, full_output_cov: bool = False, presliced=False):
        return self.K(X, X2, full_output_cov) if full_cov else self.K_diag(X, full_output_cov)


class SeparateIndependent(MultioutputKernel, Combination):
    """A different kernel per output (kernels.py:200-271): K -> [P, N, N2], K_diag -> [N, P]."""

    def __init__(self, kernels: Sequence[Kernel], name: Optional[str] = None) -> None:
        Kernel.__init__(self, name=name)
        self.kernels = list(kernels)

    @property
    def num_latent_gps(self) -> int:
        return len(self.kernels)

    @property
    def latent_kernels(self) -> Tuple[Kernel, ...]:
        return tuple(self.kernels)

    def K(self, X, X2=None, full_output_cov: bool = False):
        if full_output_cov:
            raise NotImplementedError("full_output_cov=True is outside the hot path")
        X = ops.to_device(X)
        X2d = None if X2 is None else ops.to_device(X2)
        N2 = X.shape[0] if X2d is None else X2d.shape[0]
        out = ops.empty((len(self.kernels), X.shape[0], N2), like=X)
        for p, k in enumerate(self.kernels):  # kernels.py:236-239
            ops.kbuild(compile_kernel(k, X.shape[1]), X, X2d, out=out[p])
        return out

    def K_diag(self, X, full_output_cov: bool = False):
        if full_output_cov:
            raise NotImplementedError("full_output_cov=True is outside the hot path")
        X = ops.to_device(X)
        tmp = ops.empty((len(self.kernels), X.shape[0]), like=X)
        for p, k in enumerate(self.kernels):  # kernels.py:265-271
            ops.kdiag(compile_kernel(k, X.shape[1]), X, out=tmp[p])
        return ops.transpose(tmp)  # [N, P]

    def __call__(self, X, X2=None, *, full_cov: bool = True, full_output_cov: bool = False, presliced=False):
        return self.K(X, X2, full_output_cov) if full_cov else self.K_diag(X, full_output_cov)
