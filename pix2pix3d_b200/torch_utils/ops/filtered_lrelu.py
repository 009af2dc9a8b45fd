"""Filtered leaky ReLU: bias -> upsample FIR -> lrelu(gain, slope, clamp) -> downsample FIR.

Surface of the reference's torch_utils/ops/filtered_lrelu.py:58 `filtered_lrelu`. The only caller in the reference is
StyleGAN3's alias-free synthesis layer (training/networks_stylegan3.py:357), which no pix2pix3D configuration
instantiates, so this entry point composes the sm_100a `upfirdn2d` and `bias_act` kernels exactly the way the
reference's own generic path does (filtered_lrelu.py:123-155, and its `rc = -1` CUDA fallback :225-231); a dedicated
fused kernel is not built. Differentiable to any order through the component ops.
"""
import numpy as np
import torch

from .. import misc
from . import bias_act
from . import upfirdn2d


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and 1 <= f.ndim <= 2
    return f.shape[-1], f.shape[0]


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple))
    padding = [int(v) for v in padding]
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    px0, px1, py0, py1 = padding
    return px0, px1, py0, py1


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=np.sqrt(2), slope=0.2, clamp=None,
                   flip_filter=False, impl='cuda'):
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    assert impl in ('ref', 'cuda')
    fu_w, fu_h = _get_filter_size(fu)
    fd_w, fd_h = _get_filter_size(fd)
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.dtype == x.dtype
        misc.assert_shape(b, [x.shape[1]])
    assert isinstance(up, int) and up >= 1 and isinstance(down, int) and down >= 1
    px0, px1, py0, py1 = _parse_padding(padding)
    assert gain == float(gain) and gain > 0
    assert slope == float(slope) and slope >= 0
    assert clamp is None or (clamp == float(clamp) and clamp >= 0)
    n, c, ih, iw = x.shape
    out_w = (iw * up + (px0 + px1) - (fu_w - 1) - (fd_w - 1) + (down - 1)) // down
    out_h = (ih * up + (py0 + py1) - (fu_h - 1) - (fd_h - 1) + (down - 1)) // down
    in_dtype = x.dtype
    x = bias_act.bias_act(x=x, b=b, impl=impl)
    x = upfirdn2d.upfirdn2d(x=x, f=fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter, impl=impl)
    x = bias_act.bias_act(x=x, act='lrelu', alpha=slope, gain=gain, clamp=clamp, impl=impl)
    x = upfirdn2d.upfirdn2d(x=x, f=fd, down=down, flip_filter=flip_filter, impl=impl)
    misc.assert_shape(x, [n, c, out_h, out_w])
    assert x.dtype == in_dtype
    return x
