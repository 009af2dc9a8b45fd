"""Attribute dictionaries and by-name object construction (reference dnnlib/util.py:41-58, 262-310)."""
import importlib
from typing import Any


class EasyDict(dict):
    """dict whose items are also attributes."""

    def __getattr__(self, name: str) -> Any:
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __delattr__(self, name: str) -> None:
        del self[name]


_PKG = __name__.rsplit('.', 2)[0]          # 'pix2pix3d_b200'
_MIRRORED_ROOTS = ('training', 'torch_utils', 'dnnlib', 'camera_utils')


def get_obj_by_name(name: str) -> Any:
    """Resolve 'pkg.mod.attr.sub' by trying the longest importable module prefix first. Names under the
    reference's top-level packages ('training.superresolution.X', ...) resolve inside this package."""
    if name.split('.', 1)[0] in _MIRRORED_ROOTS and not name.startswith(_PKG + '.'):
        try:
            return get_obj_by_name(f'{_PKG}.{name}')
        except ImportError:
            pass
    parts = name.split('.')
    last_err = None
    for cut in range(len(parts) - 1, 0, -1):
        mod_name = '.'.join(parts[:cut])
        try:
            obj = importlib.import_module(mod_name)
        except ModuleNotFoundError as err:
            # only swallow "this prefix is not a module"; propagate failures inside a real module
            if err.name is not None and not mod_name.startswith(err.name):
                raise
            last_err = err
            continue
        try:
            for attr in parts[cut:]:
                obj = getattr(obj, attr)
            return obj
        except AttributeError as err:
            last_err = err
    raise ImportError(f'cannot resolve {name!r}') from last_err


def call_func_by_name(*args, func_name: str = None, **kwargs) -> Any:
    assert func_name is not None
    fn = get_obj_by_name(func_name)
    assert callable(fn)
    return fn(*args, **kwargs)


def construct_class_by_name(*args, class_name: str = None, **kwargs) -> Any:
    return call_func_by_name(*args, func_name=class_name, **kwargs)
