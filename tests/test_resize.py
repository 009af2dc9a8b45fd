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


# ---------------------------------------------------------------------------------------------
# the kernel's band construction, compiled for the host (tools/probe/resize_bands_host.cu includes csrc/resize.cu)
# ---------------------------------------------------------------------------------------------
def _bands(exe, i, o, aa):
    import subprocess
    lines = subprocess.run([exe, str(i), str(o), str(aa)], capture_output=True, text=True, check=True).stdout.split('\n')
    pos, mats = 0, []
    for transposed in range(2):
        rows, cap = map(int, lines[pos].split()); pos += 1
        m = np.zeros((rows, i if transposed == 0 else o), np.float64)
        for r in range(rows):
            f = lines[pos].split(); pos += 1
            start, count = int(f[0]), int(f[1])
            assert count <= cap                                   # shared-memory band capacity chosen by make_axis
            for k in range(count):
                m[r, start + k] += float(f[2 + k])
        mats.append(m)
    return mats


def test_kernel_band_construction_on_host(tmp_path):
    """`make_band` / `make_axis` of csrc/resize.cu are __host__ __device__: build them for the CPU and compare the forward
    bands with the oracle's dense matrix and the transposed bands with its transpose, for up-, down- and odd resampling."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
    exe = str(tmp_path / 'resize_bands_host')
    r = subprocess.run([nvcc, '-std=c++17', '--expt-relaxed-constexpr', '-Wno-deprecated-gpu-targets', '-o', exe,
                        os.path.join(root, 'tools', 'probe', 'resize_bands_host.cu')], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    for aa in (1, 0):
        for i, o in [(128, 512), (512, 128), (64, 128), (100, 37), (37, 100), (1, 5), (5, 1), (7, 7), (3, 300), (300, 3), (129, 64)]:
            fwd, tr = _bands(exe, i, o, aa)
            want = O.networks.resize_matrix(i, o, bool(aa)).astype(np.float64)
            assert np.abs(fwd - want).max() < 5e-6, (aa, i, o)       # fp32 weights vs the fp64 oracle
            assert np.abs(tr - fwd.T).max() < 1e-7, (aa, i, o)
