"""Optimiser drivers of the hot path (mirrors gpflow/optimizers/__init__.py for the Scipy driver)."""
from .scipy import Scipy

__all__ = ["Scipy"]
