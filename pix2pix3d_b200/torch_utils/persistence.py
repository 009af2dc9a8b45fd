"""`persistent_class` decorator with the reference's surface (torch_utils/persistence.py:37-128).

Classes keep a record of their constructor arguments (`init_args`, `init_kwargs`) so that callers such as
training_loop.py / legacy.py can rebuild them. Embedding module source into pickles (the reference's way of
shipping code with checkpoints) is checkpoint I/O and out of scope for the hot path; objects pickle by
reference to this package instead.
"""
import copy
import sys

_decorated = set()


def persistent_class(orig_class):
    assert isinstance(orig_class, type)
    if orig_class in _decorated:
        return orig_class

    class Decorator(orig_class):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            # only the outermost decorated class of an instance records the arguments
            if getattr(self, '_init_args', None) is None or type(self) is Decorator:
                self._init_args = copy.deepcopy(args)
                self._init_kwargs = copy.deepcopy(kwargs)

        @property
        def init_args(self):
            return copy.deepcopy(self._init_args)

    # returns an EasyDict like the reference; dnnlib is imported lazily
    def _init_kwargs(self):
        from .. import dnnlib
        return dnnlib.EasyDict(copy.deepcopy(self._init_kwargs))
    Decorator.init_kwargs = property(_init_kwargs)

    Decorator.__name__ = orig_class.__name__
    Decorator.__qualname__ = orig_class.__qualname__
    Decorator.__module__ = orig_class.__module__
    Decorator.__doc__ = orig_class.__doc__
    _decorated.add(Decorator)
    return Decorator


def is_persistent(obj):
    try:
        if obj in _decorated:
            return True
    except TypeError:
        pass
    return type(obj) in _decorated


def import_hook(hook):  # accepted for API compatibility; no embedded source to rewrite
    assert callable(hook)
