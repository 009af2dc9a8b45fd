"""Fused filtered_lrelu kernel (csrc/filtered_lrelu.cu, p3d_filtered_lrelu / p3d_filtered_lrelu_act) against the four-op
reference composition (filtered_lrelu.py:123-155), the reference fixture, autograd of the composition (the gradient runs the
same kernel in sign-read mode) and the generic fallback with the in-place activation."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _kaiser(n, up):
    import scipy.signal
    return torch.as_tensor(scipy.signal.firwin(numtaps=n, cutoff=0.25 if up > 1 else 0.4, width=0.3, fs=2.0 if up == 1 else up * 1.0), dtype=torch.float32)


CASES = [
    # up, down, fu, fd, padding, clamp, flip
    (2, 2, 'f4_2d', 'f4_2d', [3, 2, 3, 2], 0.8, False),
    (2, 2, 'k12', 'k12', [10, 9, 10, 9], 256.0, False),          # StyleGAN3-style separable Kaiser filters (networks_stylegan3.py:311-318)
    (4, 2, 'k24', 'k12', [16, 17, 16, 17], None, False),
    (1, 1, None, None, 0, 1.0, False),
    (2, 1, 'k12', None, [5, 6, 5, 6], 0.5, True),
    (1, 2, None, 'k12', [5, 6, 5, 6], None, False),
    (2, 2, 'f35_2d', 'f4_2d', [4, 1, 2, 3], 0.7, True),          # asymmetric 2-D filter: flip matters
]


def _filt(kind, dev):
    from pix2pix3d_b200.torch_utils.ops import upfirdn2d
    if kind is None:
        return None
    if kind == 'f4_2d':
        return upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
    if kind == 'f35_2d':
        g = torch.Generator().manual_seed(3)
        return torch.rand(3, 5, generator=g).to(dev)
    n = int(kind[1:])
    return _kaiser(n, 2).to(dev)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
@pytest.mark.parametrize('up,down,fu,fd,padding,clamp,flip', CASES)
def test_fused_matches_composition_forward_and_backward(up, down, fu, fd, padding, clamp, flip, dtype):
    from pix2pix3d_b200 import _lib
    from pix2pix3d_b200.torch_utils.ops import filtered_lrelu as fl
    dev = torch.device('cuda')
    torch.manual_seed(0)
    x = torch.randn(2, 5, 23, 19, device=dev, dtype=dtype)
    b = torch.randn(5, device=dev, dtype=dtype)
    # fp16 operands make x + b == 0 likely somewhere; at exactly 0 the CUDA op (and the reference's kernel, `v < 0`,
    # filtered_lrelu.cu:1140) takes the positive branch while autograd of F.leaky_relu takes the slope: keep away from it
    pre = x.float() + b.float()[None, :, None, None]
    x = torch.where(pre.abs() < 1e-3, x + 0.0625, x)
    fu_t, fd_t = _filt(fu, dev), _filt(fd, dev)
    kw = dict(up=up, down=down, padding=padding, gain=1.3, slope=0.2, clamp=clamp, flip_filter=flip)
    xr, br = x.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = fl._filtered_lrelu_ref(xr, fu=fu_t, fd=fd_t, b=br, **kw)          # filters stay fp32, arithmetic in fp64
    xg, bg = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    n0 = _lib.launch_count
    y = fl.filtered_lrelu(xg, fu_t, fd_t, bg, **kw)
    assert _lib.launch_count > n0 and y.dtype == dtype and y.shape == ref.shape
    tol = 2e-5 if dtype == torch.float32 else 4e-3
    assert rel_err(y.detach().float().cpu().numpy(), ref.detach().cpu().numpy()) < tol
    gy = torch.randn(ref.shape, device=dev, dtype=torch.float64)
    gxr, gbr = torch.autograd.grad((ref * gy).sum(), [xr, br])
    gx, gb = torch.autograd.grad((y * gy.to(dtype)).sum(), [xg, bg])
    # elements whose pre-activation sits within rounding of 0 / the clamp may take the other branch in low precision
    assert rel_err(gx.float().cpu().numpy(), gxr.cpu().numpy()) < (1e-4 if dtype == torch.float32 else 3e-2)
    assert rel_err(gb.float().cpu().numpy(), gbr.cpu().numpy()) < (1e-4 if dtype == torch.float32 else 3e-2)


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def test_reference_fixture():
    from pix2pix3d_b200.torch_utils.ops import filtered_lrelu as fl
    g = load_golden('ops')
    t = lambda k: torch.from_numpy(g[k]).cuda()
    y = fl.filtered_lrelu(t('fl_x'), t('up_f4'), t('fl_fd'), t('fl_b'), up=2, down=2, padding=[3, 2, 3, 2], clamp=0.8)
    assert rel_err(y.cpu().numpy(), g['fl_up2_down2']) < 1e-5


def test_sign_tensor_contract_and_fallback_path():
    """Write mode records 1 (negative) / 2 (clamped) per up-sampled element, four per byte; the fallback composition with
    the in-place activation kernel gives the same output and the same records; read mode reproduces slope / zero from them."""
    from pix2pix3d_b200.torch_utils.ops import filtered_lrelu as fl
    from pix2pix3d_b200.torch_utils.ops import upfirdn2d
    dev = torch.device('cuda')
    torch.manual_seed(1)
    x = torch.randn(2, 3, 17, 21, device=dev)
    b = torch.randn(3, device=dev)
    f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
    up, down, pads, gain, slope, clamp = 2, 2, (3, 2, 3, 2), 1.4, 0.2, 0.9
    y, so, rc = fl._plugin_filtered_lrelu(x, f, f, b, torch.empty(0), up, down, *pads, 0, 0, gain, slope, clamp, False, True)
    assert rc == 0 and so.dtype == torch.uint8
    # generic path: bias, upsample, in-place activation (+ signs), downsample
    u = upfirdn2d.upfirdn2d(x + b[None, :, None, None], f, up=up, padding=list(pads), gain=up ** 2)
    pre = u * gain
    so2 = fl._plugin_filtered_lrelu_act_(u, torch.empty(0), 0, 0, gain, slope, clamp, True)
    y2 = upfirdn2d.upfirdn2d(u, f, down=down)
    assert rel_err(y.cpu().numpy(), y2.cpu().numpy()) < 1e-5
    # decode records over the active area and compare with the definition
    sh, swb = so.shape[2], so.shape[3]
    bits = torch.stack([(so >> (2 * k)) & 3 for k in range(4)], -1).reshape(2, 3, sh, swb * 4)
    bits2 = torch.stack([(so2 >> (2 * k)) & 3 for k in range(4)], -1).reshape(2, 3, so2.shape[2], so2.shape[3] * 4)
    aw = y.shape[3] * down - (down - 1) + 3
    want = torch.where(pre.abs() * torch.where(pre < 0, slope, 1.0) > clamp, 2, torch.where(pre < 0, 1, 0))[:, :, :sh, :aw]
    margin = ((pre.abs() - 0).abs() < 1e-5) | (((pre.abs() * torch.where(pre < 0, slope, 1.0)) - clamp).abs() < 1e-5)
    ok = (bits[:, :, :, :aw] == want) | margin[:, :, :sh, :aw]
    assert ok.all()
    ok2 = (bits2[:, :, :sh, :aw] == want) | margin[:, :, :sh, :aw]
    assert ok2.all()
    # read mode: same values as applying the recorded decision
    v = torch.randn_like(u)
    w = v.clone()
    fl._plugin_filtered_lrelu_act_(w, so2, 0, 0, 2.0, slope, float('inf'), False)
    dec = bits2[:, :, :u.shape[2], :u.shape[3]]
    expect = v * 2.0 * torch.where(dec == 1, slope, 1.0) * (dec != 2)
    assert rel_err(w.cpu().numpy(), expect.cpu().numpy()) < 1e-6


def test_unsupported_configuration_reports_minus_one_and_falls_back():
    """More taps than the kernel's tile plan admits: rc = -1 from the plugin entry, RuntimeWarning + composition from the op."""
    from pix2pix3d_b200.torch_utils.ops import filtered_lrelu as fl
    dev = torch.device('cuda')
    x = torch.randn(1, 2, 40, 40, device=dev)
    f40 = torch.ones(40, device=dev) / 40
    y, so, rc = fl._plugin_filtered_lrelu(x, f40, torch.ones(1, 1, device=dev), torch.zeros(2, device=dev), torch.empty(0), 1, 1, 20, 19, 20, 19,
                                          0, 0, 1.0, 0.2, float('inf'), False, False)
    assert rc == -1 and y is None
    with pytest.warns(RuntimeWarning):
        out = fl.filtered_lrelu(x, fu=f40, up=1, padding=[20, 19, 20, 19], gain=1.0)
    ref = fl._filtered_lrelu_ref(x, fu=f40, up=1, padding=[20, 19, 20, 19], gain=1.0)
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-5
