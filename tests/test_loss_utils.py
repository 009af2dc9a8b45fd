"""`cross_entropy2d` (reference training/loss_utils.py:4-18). CPU part: the oracle restatement against ATen's
F.cross_entropy (where the arithmetic lives); GPU part: p3d_cross_entropy2d_fwd/bwd through the mirror module."""
import numpy as np
import pytest
import torch

import p3d_oracle as O
from conftest import rel_err


def case(seed, n, c, h, w, ignore_frac=0.0):
    rng = np.random.RandomState(seed)
    x = (rng.randn(n, c, h, w) * 3).astype(np.float32)
    t = rng.randint(0, c, size=(n, h, w)).astype(np.int64)
    if ignore_frac:
        t[rng.rand(n, h, w) < ignore_frac] = -100
    wgt = (rng.rand(c) * 4 + 0.2).astype(np.float32)
    return x, t, wgt


def aten(x, t, wgt):
    xt = torch.from_numpy(x).requires_grad_(True)
    c = x.shape[1]
    flat = xt.transpose(1, 2).transpose(2, 3).contiguous().view(-1, c)       # loss_utils.py:12
    loss = torch.nn.functional.cross_entropy(flat, torch.from_numpy(t).view(-1), weight=None if wgt is None else torch.from_numpy(wgt),
                                             reduction='mean')
    loss.backward()
    return loss.item(), xt.grad.numpy()


@pytest.mark.parametrize('weighted', [False, True])
@pytest.mark.parametrize('shape,ignore', [((2, 6, 16, 16), 0.0), ((1, 19, 9, 33), 0.2), ((3, 1, 4, 4), 0.0)])
def test_oracle_matches_aten(shape, ignore, weighted):
    x, t, wgt = case(1, *shape, ignore_frac=ignore)
    wgt = wgt if weighted else None
    loss, grad = O.ops.cross_entropy2d(x, t, wgt)
    rl, rg = aten(x, t, wgt)
    assert abs(loss - rl) <= 2e-6 * max(abs(rl), 1)
    assert rel_err(grad, rg) < 1e-5


def test_mirror_cpu_path_is_the_reference_composition():
    from pix2pix3d_b200.training.loss_utils import cross_entropy2d
    x, t, wgt = case(2, 2, 6, 8, 8)
    got = cross_entropy2d(torch.from_numpy(x), torch.from_numpy(t), weight=torch.from_numpy(wgt))
    assert abs(got.item() - aten(x, t, wgt)[0]) < 1e-6
    # logits at a lower resolution than the labels are upsampled first (loss_utils.py:8-10)
    lo = torch.from_numpy(x[:, :, ::2, ::2].copy())
    ref = torch.nn.functional.interpolate(lo, size=(8, 8), mode='bilinear', align_corners=True)
    assert torch.allclose(cross_entropy2d(lo, torch.from_numpy(t)), cross_entropy2d(ref, torch.from_numpy(t)))


@pytest.mark.gpu
@pytest.mark.parametrize('weighted', [False, True])
@pytest.mark.parametrize('shape,ignore', [((2, 6, 16, 16), 0.0), ((1, 19, 9, 33), 0.2), ((3, 1, 4, 4), 0.0),
                                          ((4, 6, 512, 512), 0.0), ((4, 19, 128, 128), 0.05)])
def test_cross_entropy_kernel_matches_oracle(shape, ignore, weighted):
    from pix2pix3d_b200.training.loss_utils import cross_entropy2d
    x, t, wgt = case(3, *shape, ignore_frac=ignore)
    wgt = wgt if weighted else None
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    loss = cross_entropy2d(xt, torch.from_numpy(t).cuda(), weight=None if wgt is None else torch.from_numpy(wgt).cuda())
    (loss * 1.7).backward()
    ol, og = O.ops.cross_entropy2d(x, t, wgt)
    assert abs(loss.item() - ol) <= 2e-6 * max(abs(ol), 1)
    assert rel_err(xt.grad.cpu().numpy(), og * 1.7) < 1e-5
    # deterministic: fixed-order reduction
    loss2 = cross_entropy2d(xt.detach(), torch.from_numpy(t).cuda(), weight=None if wgt is None else torch.from_numpy(wgt).cuda())
    assert loss2.item() == loss.item()


@pytest.mark.gpu
def test_cross_entropy_all_ignored_is_nan_like_aten():
    from pix2pix3d_b200.training.loss_utils import cross_entropy2d
    x = torch.randn(1, 4, 8, 8, device='cuda')
    t = torch.full((1, 8, 8), -100, dtype=torch.int64, device='cuda')
    assert torch.isnan(cross_entropy2d(x, t))


@pytest.mark.parametrize('wtag', ['plain', 'weighted'])
def test_oracle_and_mirror_match_reference_cross_entropy2d(wtag):
    """Fixture produced by the reference's own `cross_entropy2d` (oracle/make_golden.py loss_ops)."""
    from conftest import load_golden
    from pix2pix3d_b200.training.loss_utils import cross_entropy2d
    g = load_golden('loss_ops')
    for tag in ('same', 'lowres'):
        x, t, w = g[f'ce_{tag}_x'], g[f'ce_{tag}_t'], g[f'ce_{tag}_w']
        wgt = w if wtag == 'weighted' else None
        want, want_gx = float(g[f'ce_{tag}_{wtag}_loss']), g[f'ce_{tag}_{wtag}_gx']
        xt = torch.from_numpy(x).requires_grad_(True)
        loss = cross_entropy2d(xt, torch.from_numpy(t), weight=None if wgt is None else torch.from_numpy(wgt))
        (gx,) = torch.autograd.grad(loss, xt)
        assert abs(loss.item() - want) <= 1e-6 * max(abs(want), 1)
        assert rel_err(gx.numpy(), want_gx) < 1e-6
        if tag == 'same':       # the oracle restates the per-pixel loss (the label upsample of :8-10 stays an ATen call)
            ol, og = O.ops.cross_entropy2d(x, t, wgt)
            assert abs(ol - want) <= 2e-6 * max(abs(want), 1)
            assert rel_err(og, want_gx) < 1e-5
