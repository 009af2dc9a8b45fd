"""Bilinear / anti-aliased resize (superresolution.py:315-319, dual_discriminator.py:86-102 in the reference).

CPU part: the oracle restatement against ATen's `F.interpolate` (the third-party code the arithmetic lives in, pinned by
running it). GPU part: `p3d_resize_bilinear` through `torch_utils.ops.resize` against the oracle, forward, adjoint and
second order."""
import numpy as np
import pytest
import torch

import p3d_oracle as O
from conftest import rel_err

SIZES = [((128, 128), (512, 512)), ((512, 512), (128, 128)), ((64, 64), (128, 128)), ((100, 37), (37, 100)),
         ((37, 100), (64, 21)), ((1, 5), (5, 1)), ((7, 9), (7, 9)), ((3, 300), (250, 3))]


@pytest.mark.parametrize('antialias', [True, False])
@pytest.mark.parametrize('src,dst', SIZES)
def test_oracle_resize_matches_aten(src, dst, antialias):
    rng = np.random.RandomState(src[0] * 7 + dst[1])
    if src[1] == 1 or dst[1] == 1:
        pytest.skip('ATen CPU anti-aliased kernel mishandles width-1 images')
    x = rng.randn(2, 3, *src).astype(np.float32)
    ref = torch.nn.functional.interpolate(torch.from_numpy(x), size=dst, mode='bilinear', align_corners=False, antialias=antialias)
    got = O.networks.bilinear_resize(x, dst, antialias)
    assert rel_err(got, ref.numpy()) < 1e-5
    g = rng.randn(2, 3, *dst).astype(np.float32)
    xt = torch.from_numpy(x).requires_grad_(True)
    y = torch.nn.functional.interpolate(xt, size=dst, mode='bilinear', align_corners=False, antialias=antialias)
    y.backward(torch.from_numpy(g))
    assert rel_err(O.networks.bilinear_resize_adjoint(g, src, antialias), xt.grad.numpy()) < 1e-5


def test_cpu_tensors_take_the_aten_path():
    from pix2pix3d_b200.torch_utils.ops.resize import interpolate_bilinear
    x = torch.randn(1, 2, 16, 16)
    y = interpolate_bilinear(x, (8, 8), antialias=True)
    assert torch.equal(y, torch.nn.functional.interpolate(x, size=(8, 8), mode='bilinear', align_corners=False, antialias=True))


@pytest.mark.gpu
@pytest.mark.parametrize('antialias', [True, False])
@pytest.mark.parametrize('src,dst', SIZES + [((1, 5), (5, 1))[::-1], ((5, 1), (1, 7))])
def test_resize_kernel_matches_oracle(src, dst, antialias):
    from pix2pix3d_b200.torch_utils.ops.resize import interpolate_bilinear
    rng = np.random.RandomState(src[1] * 5 + dst[0])
    x = rng.randn(2, 5, *src).astype(np.float32)
    g = rng.randn(2, 5, *dst).astype(np.float32)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    y = interpolate_bilinear(xt, dst, antialias=antialias)
    assert y.shape == (2, 5, *dst)
    assert rel_err(y.detach().cpu().numpy(), O.networks.bilinear_resize(x, dst, antialias)) < 1e-5
    y.backward(torch.from_numpy(g).cuda())
    assert rel_err(xt.grad.cpu().numpy(), O.networks.bilinear_resize_adjoint(g, src, antialias)) < 1e-5
    # half precision storage, fp32 accumulation
    yh = interpolate_bilinear(torch.from_numpy(x).cuda().half(), dst, antialias=antialias)
    assert yh.dtype == torch.float16
    assert rel_err(yh.float().cpu().numpy(), O.networks.bilinear_resize(x.astype(np.float16).astype(np.float32), dst, antialias)) < 2e-3


@pytest.mark.gpu
def test_resize_kernel_agrees_with_aten_cuda_at_training_sizes():
    """The two shapes of config 5: raw render 128 -> 512 for D (dual_discriminator.py:157-160) and real image 512 -> 128 for
    the loss (loss.py `filtered_resizing` of the real image)."""
    from pix2pix3d_b200.torch_utils.ops.resize import interpolate_bilinear
    gen = torch.Generator().manual_seed(3)
    for src, dst, c in ((128, 512, 6), (512, 128, 3)):
        x = torch.randn(4, c, src, src, generator=gen).cuda()
        ref = torch.nn.functional.interpolate(x, size=(dst, dst), mode='bilinear', align_corners=False, antialias=True)
        got = interpolate_bilinear(x, (dst, dst), antialias=True)
        assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 1e-5


@pytest.mark.gpu
def test_resize_kernel_gradients_to_second_order():
    from pix2pix3d_b200.torch_utils.ops.resize import interpolate_bilinear
    x = torch.randn(1, 2, 9, 6, dtype=torch.float64, device='cuda', requires_grad=True)
    for aa in (True, False):
        fn = lambda t: interpolate_bilinear(t, (4, 11), antialias=aa)
        assert torch.autograd.gradcheck(fn, (x,), eps=1e-6, atol=1e-5)
        assert torch.autograd.gradgradcheck(fn, (x,), eps=1e-6, atol=1e-5)
    # the R1 pattern: gradient of a scalar w.r.t. the input of the resize, differentiated again
    xr = torch.randn(2, 3, 32, 32, device='cuda', requires_grad=True)
    w = torch.randn(2, 3, 64, 64, device='cuda', requires_grad=True)
    y = interpolate_bilinear(xr, (64, 64)) * w
    (gx,) = torch.autograd.grad(y.sum(), xr, create_graph=True)
    gx.square().sum().backward()
    ref_gx = O.networks.bilinear_resize_adjoint(w.detach().cpu().numpy(), (32, 32))
    assert rel_err(gx.detach().cpu().numpy(), ref_gx) < 1e-5
    # d/dw sum((A^T w)^2) = 2 A A^T w
    ref_gw = 2 * O.networks.bilinear_resize(ref_gx, (64, 64))
    assert rel_err(w.grad.cpu().numpy(), ref_gw) < 1e-5


# ---------------------------------------------------------------------------------------------
# fixture produced by the reference's own `filtered_resizing` (oracle/make_golden.py loss_ops)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize('tag,dst', [('up', 64), ('down', 16), ('odd', 37)])
def test_oracle_and_mirror_match_reference_filtered_resizing(tag, dst):
    from conftest import load_golden
    from pix2pix3d_b200.torch_utils.ops import upfirdn2d
    from pix2pix3d_b200.training.dual_discriminator import filtered_resizing
    g = load_golden('loss_ops')
    x, gy = g[f'fr_{tag}_x'], g[f'fr_{tag}_gy']
    # oracle restatement: the two pure-interpolation modes, forward and input gradient
    for mode, aa in (('antialiased', True), ('none', False)):
        assert rel_err(O.networks.bilinear_resize(x, (dst, dst), aa), g[f'fr_{tag}_{mode}_y']) < 1e-5
        assert rel_err(O.networks.bilinear_resize_adjoint(gy, x.shape[2:], aa), g[f'fr_{tag}_{mode}_gx']) < 1e-5
    # mirror module on CPU tensors: every filter mode of the reference, bit for bit (same ATen calls)
    f4 = upfirdn2d.setup_filter([1, 3, 3, 1])
    for mode in ('antialiased', 'none', 0.3) + (('classic',) if tag == 'up' else ()):
        xt = torch.from_numpy(x).requires_grad_(True)
        y = filtered_resizing(xt, size=dst, f=f4, filter_mode=mode)
        (gx,) = torch.autograd.grad((y * torch.from_numpy(gy)).sum(), xt)
        assert rel_err(y.detach().numpy(), g[f'fr_{tag}_{mode}_y']) < 1e-6
        assert rel_err(gx.numpy(), g[f'fr_{tag}_{mode}_gx']) < 1e-6
