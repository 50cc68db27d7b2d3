"""Type-based multiple dispatch — the reference's plugin mechanism
(mirrors gpflow/utilities/multipledispatch.py:29-85; implemented here without the external
`multipledispatch` package, which is not part of this image)."""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple, Type


class Dispatcher:
    def __init__(self, name: str) -> None:
        self.name = name
        self.funcs: Dict[Tuple[type, ...], Callable[..., Any]] = {}
        self._cache: Dict[Tuple[type, ...], Callable[..., Any]] = {}

    def register(self, *types: Any) -> Callable[[Callable[..., Any]], Callable[..., Any]]:
        """`types` entries may be a type or a tuple of types (union), as in multipledispatch."""

        def deco(fn: Callable[..., Any]) -> Callable[..., Any]:
            def expand(ts: Tuple[Any, ...]) -> List[Tuple[type, ...]]:
                if not ts:
                    return [()]
                head = ts[0] if isinstance(ts[0], tuple) else (ts[0],)
                return [(h,) + rest for h in head for rest in expand(ts[1:])]

            for sig in expand(tuple(types)):
                self.funcs[sig] = fn
            self._cache.clear()
            return fn

        return deco

    def dispatch(self, *types: type) -> Optional[Callable[..., Any]]:
        if types in self._cache:
            return self._cache[types]
        best, best_score = None, None
        for sig, fn in self.funcs.items():
            if len(sig) != len(types):
                continue
            score = []
            for t, s in zip(types, sig):
                if s is object:
                    score.append(len(t.__mro__))
                elif issubclass(t, s):
                    score.append(t.__mro__.index(s))
                else:
                    break
            else:
                tup = tuple(score)
                if best_score is None or tup < best_score:
                    best, best_score = fn, tup
        if best is not None:
            self._cache[types] = best
        return best

    def dispatch_or_raise(self, *types: type) -> Callable[..., Any]:  # multipledispatch.py:50-63
        fn = self.dispatch(*types)
        if fn is None:
            raise NotImplementedError(
                f"Could not find signature for {self.name}: <{', '.join(t.__name__ for t in types)}>"
            )
        return fn

    def __call__(self, *args: Any, **kwargs: Any) -> Any:
        n = max((len(s) for s in self.funcs), default=0)
        types = tuple(type(a) for a in args[:n])
        fn = self.dispatch(*types)
        while fn is None and len(types) > 1:
            types = types[:-1]
            fn = self.dispatch(*types)
        if fn is None:
            raise NotImplementedError(
                f"Could not find signature for {self.name}: <{', '.join(type(a).__name__ for a in args)}>"
            )
        return fn(*args, **kwargs)
