"""Fused multiply-add `a * b + c` with broadcast-aware gradients (reference torch_utils/ops/fma.py:17-62)."""
import torch


def fma(a, b, c):
    return _Fma.apply(a, b, c)


def _sum_to_shape(x, shape):
    """Reduce `x` over the dimensions that were broadcast from `shape`."""
    extra = x.ndim - len(shape)
    assert extra >= 0
    dims = [i for i in range(x.ndim) if i < extra or (x.shape[i] > 1 and shape[i - extra] == 1)]
    if dims:
        x = x.sum(dim=dims, keepdim=True)
    if extra:
        x = x.reshape(-1, *x.shape[extra + 1:])
    assert tuple(x.shape) == tuple(shape)
    return x


class _Fma(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):
        out = torch.addcmul(c, a, b)
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b = ctx.saved_tensors
        da = db = dc = None
        if ctx.needs_input_grad[0]:
            da = _sum_to_shape(fma(dout, b, torch.zeros([], dtype=dout.dtype, device=dout.device)), a.shape)
        if ctx.needs_input_grad[1]:
            db = _sum_to_shape(fma(dout, a, torch.zeros([], dtype=dout.dtype, device=dout.device)), b.shape)
        if ctx.needs_input_grad[2]:
            dc = _sum_to_shape(dout, ctx.c_shape)
        return da, db, dc
