"""Inducing variables (mirrors gpflow/inducing_variables/inducing_variables.py:27-94)."""
from __future__ import annotations

import abc
from typing import Any, Optional, Tuple

from .base import Module, Parameter


class InducingVariables(Module, metaclass=abc.ABCMeta):
    @property
    @abc.abstractmethod
    def num_inducing(self) -> int:
        raise NotImplementedError

    @property
    @abc.abstractmethod
    def shape(self) -> Tuple[Optional[int], ...]:
        raise NotImplementedError


class InducingPointsBase(InducingVariables):
    def __init__(self, Z: Any, name: Optional[str] = None):
        self.name = name
        if not isinstance(Z, Parameter):
            Z = Parameter(Z)
        if len(Z.shape) != 2:
            raise ValueError("Z must have shape [M, D]")
        self.Z = Z

    @property
    def num_inducing(self) -> int:
        return self.Z.shape[0]

    @property
    def shape(self) -> Tuple[Optional[int], ...]:
        return (self.Z.shape[0], self.Z.shape[1], 1)


class InducingPoints(InducingPointsBase):
    """Real-space inducing points."""


def inducingpoint_wrapper(inducing_variable: Any) -> InducingVariables:
    """gpflow/models/util.py:31-38."""
    if not isinstance(inducing_variable, InducingVariables):
        inducing_variable = InducingPoints(inducing_variable)
    return inducing_variable
