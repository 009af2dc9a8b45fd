"""conv2d_gradfix autograd nodes (reference torch_utils/ops/conv2d_gradfix.py:68-197) against plain autograd:
first and second order, conv2d and conv_transpose2d, and the R1 pattern (grad of |dy/dx|^2 w.r.t. w)."""
import pytest
import torch

from pix2pix3d_b200.torch_utils.ops import conv2d_gradfix as cg


def _ops(transpose, wshape, stride, padding, output_padding, groups):
    custom = cg._make_op(transpose, tuple(wshape), (stride, stride), (padding, padding), (output_padding, output_padding), (1, 1), groups)
    if transpose:
        plain = lambda x, w, b: torch.nn.functional.conv_transpose2d(x, w, b, stride=stride, padding=padding, output_padding=output_padding, groups=groups)
    else:
        plain = lambda x, w, b: torch.nn.functional.conv2d(x, w, b, stride=stride, padding=padding, groups=groups)
    return custom.apply, plain


CASES = [
    # transpose, x shape, w shape, stride, padding, output_padding, groups
    (False, (2, 4, 9, 9), (6, 4, 3, 3), 1, 1, 0, 1),
    (False, (2, 4, 10, 10), (6, 4, 3, 3), 2, 0, 0, 1),
    (False, (2, 4, 8, 8), (6, 2, 3, 3), 1, 1, 0, 2),
    (False, (2, 4, 8, 8), (6, 4, 1, 1), 1, 0, 0, 1),
    (True, (2, 4, 5, 5), (4, 6, 3, 3), 2, 0, 0, 1),
    (True, (2, 4, 5, 5), (4, 3, 3, 3), 2, 1, 1, 2),
    (True, (2, 4, 6, 6), (4, 6, 1, 1), 1, 0, 0, 1),
]


@pytest.mark.parametrize('transpose,xs,wshape,stride,padding,opad,groups', CASES)
def test_first_and_second_order_grads_match_autograd(transpose, xs, wshape, stride, padding, opad, groups):
    torch.manual_seed(0)
    custom, plain = _ops(transpose, wshape, stride, padding, opad, groups)
    cout = wshape[1] * groups if transpose else wshape[0]
    res = []
    for f in (custom, plain):
        x = torch.randn(xs, dtype=torch.float64, requires_grad=True)
        torch.manual_seed(1)
        w = torch.randn(wshape, dtype=torch.float64, requires_grad=True)
        b = torch.randn(cout, dtype=torch.float64, requires_grad=True)
        x.data.copy_(torch.randn(xs, dtype=torch.float64, generator=torch.Generator().manual_seed(2)))
        y = f(x, w, b)
        gy = torch.randn(y.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
        dx, dw, db = torch.autograd.grad((y * gy).sum(), [x, w, b], create_graph=True)
        # R1-like second order: d/d{x,w} of |dy/dx|^2 + |dy/dw|^2
        pen = dx.square().sum() + dw.square().sum()
        d2x, d2w = torch.autograd.grad(pen, [x, w], allow_unused=True)
        res.append((y, dx, dw, db, d2x, d2w))
    for a, r in zip(res[0], res[1]):
        if r is None:
            assert a is None or float(a.abs().max()) == 0
            continue
        torch.testing.assert_close(a, r, rtol=1e-10, atol=1e-10)


def test_r1_pattern_with_weight_gradients_disabled():
    """loss.py:873-879: grad of logits w.r.t. the image under no_weight_gradients(), then backward of the penalty
    into the weights -- the dx node of a plain conv2d is the transposed op, whose weight gradient must work."""
    torch.manual_seed(0)
    custom, plain = _ops(False, (5, 3, 3, 3), 1, 1, 0, 1)
    out = []
    for f in (custom, plain):
        img = torch.randn(2, 3, 8, 8, dtype=torch.float64, generator=torch.Generator().manual_seed(5)).requires_grad_(True)
        w = torch.randn(5, 3, 3, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(6)).requires_grad_(True)
        logits = f(img, w, None).tanh().sum()
        with cg.no_weight_gradients():
            (g,) = torch.autograd.grad(logits, [img], create_graph=True)
        pen = g.square().sum()
        (dw,) = torch.autograd.grad(pen, [w])
        out.append(dw)
    torch.testing.assert_close(out[0], out[1], rtol=1e-10, atol=1e-10)


def test_no_weight_gradients_skips_dw():
    custom, _ = _ops(False, (5, 3, 3, 3), 1, 1, 0, 1)
    x = torch.randn(1, 3, 6, 6, requires_grad=True)
    w = torch.randn(5, 3, 3, 3, requires_grad=True)
    with cg.no_weight_gradients():
        custom(x, w, None).sum().backward()
    assert w.grad is None and x.grad is not None


def test_native_conv_dispatch_is_cuda_only_and_scoped():
    """native_conv.applies: never for CPU tensors, never outside first_order(); the generator classes enter the region in
    mapping / synthesis / sample / sample_mixed and leave it again."""
    from pix2pix3d_b200.torch_utils.ops import native_conv
    import pix2pix3d_b200.training.triplane_cond as tc
    x = torch.randn(1, 8, 16, 16, requires_grad=True)
    w = torch.randn(8, 8, 3, 3, requires_grad=True)
    with native_conv.first_order():
        assert native_conv._depth == 1
        assert not native_conv.applies(x, w, None, (1, 1), (1, 1), (1, 1), 1)
        y = cg.conv2d(x, w, padding=1)
    assert native_conv._depth == 0 and y.shape == (1, 8, 16, 16)
    for cls in (tc.TriPlaneGenerator, tc.TriPlaneSemanticGenerator, tc.TriPlaneSemanticEntangleGenerator,
                tc.TriPlaneSemanticEntangleGenerator_withBG):
        for name in ('mapping', 'synthesis', 'sample', 'sample_mixed'):
            assert getattr(getattr(cls, name), '_p3d_first_order', False), (cls.__name__, name)
    import pix2pix3d_b200.training.dual_discriminator as dd
    assert not getattr(dd.DualDiscriminator.forward, '_p3d_first_order', False)      # R1 differentiates D twice
