"""Scipy optimiser driver (mirrors gpflow/optimizers/scipy.py:42-331).

The reference wraps `scipy.optimize.minimize`: it packs the unconstrained variables into one flat float64 vector
(`pack_tensors`, :322-325), evaluates loss and gradients with a tf.GradientTape (`_compute_loss_and_gradients`, :300-316)
and unpacks the optimiser's iterate back into the variables (`assign_tensors`, :333-337).  Here the gradients come from
the device backward pass of the model (`training_loss_and_gradients`, csrc/grad.cu) instead of autodiff; the packing
contract is the same."""
from __future__ import annotations

from typing import Any, Callable, Mapping, Optional, Sequence

import numpy as np

from ..base import Parameter


class Scipy:
    def __init__(self, compile_cache_size: int = 2) -> None:
        self.compile_cache_size = compile_cache_size  # kept for signature parity; nothing is traced here

    def minimize(self, closure: Callable[[], Any], variables: Sequence[Parameter], method: Optional[str] = "L-BFGS-B",
                 step_callback: Optional[Callable[..., None]] = None, compile: bool = True,
                 allow_unused_variables: bool = False, tf_fun_args: Optional[Mapping[str, Any]] = None,
                 track_loss_history: bool = False, **scipy_kwargs: Any):
        """scipy.py:78-228.  `closure` is `model.training_loss_closure()`; it must expose `value_and_gradients(variables)`
        (models with a device backward pass do)."""
        import scipy.optimize

        if not callable(closure):
            raise TypeError("The 'closure' argument is expected to be a callable object.")  # scipy.py:131-134
        variables = tuple(variables)
        if not all(isinstance(v, Parameter) for v in variables):
            raise TypeError("The 'variables' argument is expected to only contain Parameter instances")
        if not hasattr(closure, "value_and_gradients"):
            raise NotImplementedError("this closure has no device backward pass (value_and_gradients)")
        if tf_fun_args:
            raise ValueError("`tf_fun_args` should only be set when `compile` is True.") if not compile else None
        x0 = self.initial_parameters(variables)
        history = []

        def fun(x: np.ndarray):
            self.assign_tensors(variables, self.unpack_tensors(variables, x))
            loss, grads = closure.value_and_gradients(variables)
            if track_loss_history:
                history.append(float(loss))
            return float(loss), self.pack_tensors(grads)

        cb = None
        if step_callback is not None:  # scipy.py:339-352
            step = [0]

            def cb(x: np.ndarray) -> None:
                values = self.unpack_tensors(variables, x)
                step_callback(step[0], variables, values)
                step[0] += 1

        opt = scipy.optimize.minimize(fun, x0, jac=True, method=method, callback=cb, **scipy_kwargs)
        values = self.unpack_tensors(variables, opt.x)
        self.assign_tensors(variables, values)  # scipy.py:221-223
        if track_loss_history:
            opt["loss_history"] = history
        return opt

    @classmethod
    def initial_parameters(cls, variables: Sequence[Parameter]) -> np.ndarray:  # scipy.py:230-232
        return cls.pack_tensors([v.unconstrained_variable for v in variables])

    @staticmethod
    def pack_tensors(tensors: Sequence[Any]) -> np.ndarray:  # scipy.py:322-325
        return np.concatenate([np.asarray(t, dtype=np.float64).reshape(-1) for t in tensors]) if tensors else np.zeros(0)

    @staticmethod
    def unpack_tensors(to_tensors: Sequence[Parameter], from_vector: np.ndarray):  # scipy.py:327-331
        s, values = 0, []
        for p in to_tensors:
            n = int(np.prod(p.shape)) if p.shape else 1
            values.append(np.asarray(from_vector[s:s + n], dtype=np.float64).reshape(p.shape))
            s += n
        return values

    @staticmethod
    def assign_tensors(to_tensors: Sequence[Parameter], values: Sequence[np.ndarray]) -> None:  # scipy.py:333-337
        if len(to_tensors) != len(values):
            raise ValueError("to_tensors and values should have same length")
        for p, v in zip(to_tensors, values):
            p.assign_unconstrained(v)
