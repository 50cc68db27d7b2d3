from .multipledispatch import Dispatcher
from .misc import set_trainable, to_default_float

__all__ = ["Dispatcher", "set_trainable", "to_default_float"]
