"""gpflow/utilities/misc.py:47-74 subset."""
from __future__ import annotations

from typing import Any, Iterable, Union

import numpy as np

from .. import config
from ..base import Module, Parameter


def to_default_float(x: Any):
    if hasattr(x, "detach"):
        from .. import ops

        return ops.to_device(x)
    return np.asarray(x, dtype=config.default_float())


def set_trainable(model: Union[Module, Parameter, Iterable[Any]], flag: bool) -> None:
    if isinstance(model, Parameter):
        model.trainable = flag
        return
    if isinstance(model, Module):
        for p in model.parameters:
            p.trainable = flag
        return
    for m in model:
        set_trainable(m, flag)
