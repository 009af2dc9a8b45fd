"""Small helpers the hot-path modules call (reference torch_utils/misc.py:84-107, 150-176)."""
import contextlib
import re
import warnings

import torch


def assert_shape(tensor, ref_shape):
    """Check rank and every non-None entry of `ref_shape` (reference torch_utils/misc.py:84)."""
    if tensor.ndim != len(ref_shape):
        raise AssertionError(f'Wrong number of dimensions: got {tensor.ndim}, expected {len(ref_shape)}')
    for idx, (size, ref) in enumerate(zip(tensor.shape, ref_shape)):
        if ref is None:
            continue
        if int(size) != int(ref):
            raise AssertionError(f'Wrong size for dimension {idx}: got {int(size)}, expected {int(ref)}')


@contextlib.contextmanager
def suppress_tracer_warnings():
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', category=torch.jit.TracerWarning)
        yield


def profiled_function(fn):
    """Wrap `fn` in a profiler scope named after it."""
    def wrapped(*args, **kwargs):
        with torch.autograd.profiler.record_function(fn.__name__):
            return fn(*args, **kwargs)
    wrapped.__name__ = fn.__name__
    wrapped.__doc__ = fn.__doc__
    return wrapped


def params_and_buffers(module):
    assert isinstance(module, torch.nn.Module)
    return list(module.parameters()) + list(module.buffers())


def named_params_and_buffers(module):
    assert isinstance(module, torch.nn.Module)
    return list(module.named_parameters()) + list(module.named_buffers())


def copy_params_and_buffers(src_module, dst_module, require_all=False, allow_mismatch=False):
    """Name-matched tensor copy between modules (reference torch_utils/misc.py:157-176)."""
    src = dict(named_params_and_buffers(src_module))
    for name, tensor in named_params_and_buffers(dst_module):
        # pix2pix3D addition (:163-165): `superresolution_semantic.*` is initialised from `superresolution.*` when the source
        # (an EG3D checkpoint) has no semantic branch
        name_src = name if name in src else name.replace('_semantic', '')
        if name_src not in src:
            if require_all:
                raise AssertionError(f'{name} missing in source module')
            print(f'Warning: {name} not found in source module')
            continue
        s = src[name_src].detach()
        if s.shape != tensor.shape and allow_mismatch:
            print(f'Warning: {name_src} shape mismatch: {tuple(s.shape)} vs {tuple(tensor.shape)}')
            continue
        tensor.copy_(s).requires_grad_(tensor.requires_grad)
