"""Global configuration (mirrors gpflow/config/__config__.py:87-104,181-215,263-384).

`default_float()` decides the dtype of every Parameter and data conversion, `default_jitter()`
enters every Kuu; both are part of numerical parity with the reference."""
from __future__ import annotations

import contextlib
import os
from dataclasses import dataclass, replace
from typing import Iterator, Optional

import numpy as np


@dataclass(frozen=True)
class Config:
    int: type = np.int32
    float: type = np.float64
    jitter: float = 1e-6
    positive_bijector: str = "softplus"
    positive_minimum: float = 0.0
    likelihood_positive_minimum: float = 1e-6
    summary_fmt: Optional[str] = None


def _from_env() -> Config:
    kw = {}
    f = os.environ.get("GPFLOW_FLOAT")
    if f:
        kw["float"] = {"float32": np.float32, "float64": np.float64}[f]
    j = os.environ.get("GPFLOW_JITTER")
    if j:
        kw["jitter"] = float(j)
    b = os.environ.get("GPFLOW_POSITIVE_BIJECTOR")
    if b:
        kw["positive_bijector"] = b
    return Config(**kw)


__config = _from_env()


def config() -> Config:
    return __config


def set_config(new: Config) -> None:
    global __config
    __config = new


def default_int() -> type:
    return __config.int


def default_float() -> type:
    return __config.float


def default_jitter() -> float:
    return __config.jitter


def default_positive_bijector() -> str:
    return __config.positive_bijector


def default_positive_minimum() -> float:
    return __config.positive_minimum


def default_likelihood_positive_minimum() -> float:
    return __config.likelihood_positive_minimum


def set_default_float(value_type: type) -> None:
    vt = np.dtype(value_type).type
    if vt not in (np.float32, np.float64):
        raise TypeError(f"{value_type} is not a supported float type (float32 / float64)")
    set_config(replace(__config, float=vt))


def set_default_jitter(value: float) -> None:
    if not isinstance(value, (float, int)) or value < 0:
        raise ValueError("Expected a non-negative float for the jitter")
    set_config(replace(__config, jitter=float(value)))


def set_default_positive_bijector(value: str) -> None:
    if value.lower() not in ("exp", "softplus"):
        raise ValueError(f"`{value}` not in set of valid bijectors: ['exp', 'softplus']")
    set_config(replace(__config, positive_bijector=value.lower()))


def set_default_positive_minimum(value: float) -> None:
    if value < 0:
        raise ValueError("Positive minimum must be non-negative")
    set_config(replace(__config, positive_minimum=float(value)))


def set_default_likelihood_positive_minimum(value: float) -> None:
    if value < 0:
        raise ValueError("Likelihood positive minimum must be non-negative")
    set_config(replace(__config, likelihood_positive_minimum=float(value)))


@contextlib.contextmanager
def as_context(temporary_config: Optional[Config] = None) -> Iterator[None]:
    current = config()
    set_config(temporary_config if temporary_config is not None else current)
    try:
        yield
    finally:
        set_config(current)
