"""Shared test helpers: build the SAME kernel expression for the oracle and the product from a spec."""
from __future__ import annotations

import numpy as np

from oracle import gp_oracle as O

SPECS = {
    "rbf": ("RBF", dict(variance=1.3, lengthscales=0.9)),
    "rbf_ard": ("RBF", dict(variance=0.7, lengthscales="ard")),
    "m12": ("Matern12", dict(variance=1.1, lengthscales=1.7)),
    "m32": ("Matern32", dict(variance=0.8, lengthscales=1.2)),
    "m52": ("Matern52", dict(variance=1.5, lengthscales=2.1)),
    "rq": ("RationalQuadratic", dict(variance=1.2, lengthscales=1.4, alpha=0.7)),
    "exp": ("Exponential", dict(variance=0.9, lengthscales=1.1)),
    "lin": ("Linear", dict(variance=0.6)),
    "lin_ard": ("Linear", dict(variance="ard")),
    "poly": ("Polynomial", dict(degree=3.0, variance=0.35, offset=0.8)),
    "poly_ard": ("Polynomial", dict(degree=2.0, variance="ard", offset=1.3)),
    "const": ("Constant", dict(variance=0.4)),
    "white": ("White", dict(variance=0.3)),
}


def _leaf(mod, name, kw, D, active_dims, rng):
    kw = dict(kw)
    nd = D if active_dims is None else len(active_dims)
    for k, v in list(kw.items()):
        if isinstance(v, str) and v == "ard":
            kw[k] = 0.5 + rng.random(nd)
    cls = getattr(mod, "SquaredExponential" if name == "RBF" else name)
    if active_dims is not None:
        kw["active_dims"] = active_dims
    return cls(**kw)


def build(expr, D, mods, seed=0):
    """expr: nested tuples ("sum", a, b, ...) / ("prod", a, b, ...) / leaf key or (leaf key, active_dims).
    Returns one kernel object per module in `mods` (same random ARD parameters)."""
    outs = []
    for mod in mods:
        rng = np.random.default_rng(seed)

        def rec(e):
            if isinstance(e, tuple) and e[0] in ("sum", "prod"):
                kids = [rec(c) for c in e[1:]]
                return (mod.Sum if e[0] == "sum" else mod.Product)(kids)
            key, ad = (e, None) if isinstance(e, str) else e
            name, kw = SPECS[key]
            return _leaf(mod, name, kw, D, ad, rng)

        outs.append(rec(expr))
    return outs


def tol(dtype):
    return dict(rtol=1e-11, atol=1e-12) if np.dtype(dtype) == np.float64 else dict(rtol=2e-5, atol=2e-5)


def to_np(t):
    return t.detach().cpu().numpy()


def _has_nonsmooth(expr):
    if isinstance(expr, str):
        return expr in ("m12", "exp")
    if isinstance(expr, tuple) and expr and expr[0] in ("sum", "prod"):
        return any(_has_nonsmooth(e) for e in expr[1:])
    if isinstance(expr, tuple):
        return _has_nonsmooth(expr[0])
    return False


def tol_for(expr, dtype):
    """Matern12 / Exponential are sqrt-like at r=0: the reference's norm-expansion distance
    (gpflow/utilities/ops.py:109-111) leaves rounding noise eps*|x|^2 in r^2, hence sqrt(eps) noise in K
    near the diagonal — in the reference itself.  Two correct evaluations therefore agree only to
    ~sqrt(eps) there; everywhere else the tight tolerance applies."""
    if _has_nonsmooth(expr):
        return dict(rtol=1e-7, atol=1e-7) if np.dtype(dtype) == np.float64 else dict(rtol=5e-3, atol=5e-3)
    return tol(dtype)
