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


class MultioutputInducingVariables(InducingVariables):
    """gpflow/inducing_variables/multioutput/inducing_variables.py:22-35."""

    @property
    @abc.abstractmethod
    def inducing_variables(self) -> Tuple[InducingVariables, ...]:
        raise NotImplementedError


class FallbackSharedIndependentInducingVariables(MultioutputInducingVariables):
    """One set of inducing variables shared by all latent GPs (multioutput/inducing_variables.py:38-95)."""

    def __init__(self, inducing_variable: InducingVariables):
        self.inducing_variable = inducingpoint_wrapper(inducing_variable)

    @property
    def num_inducing(self) -> int:
        return self.inducing_variable.num_inducing

    @property
    def inducing_variables(self) -> Tuple[InducingVariables, ...]:
        return (self.inducing_variable,)

    @property
    def shape(self) -> Tuple[Optional[int], ...]:
        inner = self.inducing_variable.shape
        return inner[:-1] + (None,)


class FallbackSeparateIndependentInducingVariables(MultioutputInducingVariables):
    """One set of inducing variables per latent GP (multioutput/inducing_variables.py:98-166)."""

    def __init__(self, inducing_variable_list):
        self.inducing_variable_list = [inducingpoint_wrapper(iv) for iv in inducing_variable_list]

    @property
    def num_inducing(self) -> int:  # :148-151 (they must agree)
        return self.inducing_variable_list[0].num_inducing

    @property
    def inducing_variables(self) -> Tuple[InducingVariables, ...]:
        return tuple(self.inducing_variable_list)

    @property
    def shape(self) -> Tuple[Optional[int], ...]:
        inner = self.inducing_variable_list[0].shape
        return inner[:-1] + (len(self.inducing_variable_list),)


class SharedIndependentInducingVariables(FallbackSharedIndependentInducingVariables):
    """multioutput/inducing_variables.py:169-175."""


class SeparateIndependentInducingVariables(FallbackSeparateIndependentInducingVariables):
    """multioutput/inducing_variables.py:178-184."""


def inducingpoint_wrapper(inducing_variable: Any) -> InducingVariables:
    """gpflow/models/util.py:31-38."""
    if not isinstance(inducing_variable, InducingVariables):
        inducing_variable = InducingPoints(inducing_variable)
    return inducing_variable
