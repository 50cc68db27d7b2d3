"""gpflow.kernels surface for the hot path (RBF / Matern / Linear / White / Constant, Sum / Product,
independent multi-output wrappers)."""
from .base import Combination, Kernel, Product, ReducingCombination, Sum, compile_kernel
from .linears import Linear, Polynomial
from .multioutput import MultioutputKernel, SeparateIndependent, SharedIndependent
from .statics import Bias, Constant, Static, White
from .stationaries import (
    Exponential,
    IsotropicStationary,
    Matern12,
    Matern32,
    Matern52,
    RationalQuadratic,
    SquaredExponential,
    Stationary,
)

RBF = SquaredExponential

__all__ = [
    "Bias", "Combination", "Constant", "Exponential", "IsotropicStationary", "Kernel", "Linear", "Matern12", "Polynomial",
    "Matern32", "Matern52", "MultioutputKernel", "Product", "RBF", "RationalQuadratic", "ReducingCombination",
    "SeparateIndependent", "SharedIndependent", "SquaredExponential", "Static", "Stationary", "Sum", "White",
    "compile_kernel",
]
