"""pix2pix3d_b200 -- sm_100a implementation of pix2pix3D's volumetric-rendering + StyleGAN2-op hot path.

Layout
  csrc/            CUDA kernels and the C-ABI (include/p3d.h) -> libp3d.so (built by `pix2pix3d_b200.build`)
  _lib.py          ctypes binding of the C-ABI
  native.py        tensor-level wrappers (allocate outputs, pass raw pointers)
  torch_utils/, training/, dnnlib/
                   host-side mirror of the reference's operator surface (same import paths below the package
                   root, same class / function names and signatures)

`install()` makes the mirror answer to the reference's top-level module names (`training.*`, `torch_utils.*`,
`dnnlib`, `camera_utils`, `legacy`), so reference callers (applications/generate_samples.py, train.py) and pickles that
refer to those paths resolve to this package.
"""
import importlib
import importlib.abc
import importlib.util
import sys

__version__ = '0.1.0'

_ALIASED_ROOTS = ('training', 'torch_utils', 'dnnlib', 'camera_utils', 'legacy')


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Resolves `training.x.y` to the already-imported (or importable) `pix2pix3d_b200.training.x.y`, returning
    the very same module object so that classes are not duplicated under two names."""

    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split('.', 1)[0]
        if root not in _ALIASED_ROOTS:
            return None
        real = f'{__name__}.{fullname}'
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except (ImportError, ValueError):
            return None
        return importlib.util.spec_from_loader(fullname, self, origin=real)

    def create_module(self, spec):
        return importlib.import_module(spec.origin)

    def exec_module(self, module):
        pass


_finder = None


def install():
    """Register the alias finder (idempotent). Call before importing `training`, `torch_utils` or `dnnlib`."""
    global _finder
    if _finder is None:
        _finder = _AliasFinder()
        sys.meta_path.insert(0, _finder)
    return _finder


def uninstall():
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        _finder = None
    for name in list(sys.modules):
        if name.split('.', 1)[0] in _ALIASED_ROOTS and getattr(sys.modules[name], '__name__', '').startswith(__name__ + '.'):
            del sys.modules[name]
