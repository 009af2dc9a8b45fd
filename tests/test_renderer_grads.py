"""First-order gradients of the differentiable renderer stages (tri-plane lookup, ray marcher).

CPU part: the oracle's hand-derived backward formulas against torch autograd of the reference's op composition (the
mirror's torch formulation, which tests/test_mirror_cpu.py ties to the reference). GPU part: the CUDA backward kernels
(p3d_sample_from_planes_bwd, p3d_ray_march_bwd) against the oracle and against ATen autograd on the GPU."""
import numpy as np
import pytest
import torch

import p3d_oracle as O
from conftest import rel_err


def march_case(seed, b=2, r=9, s=13, c=5, empty=True):
    rng = np.random.RandomState(seed)
    depths = np.sort(rng.rand(b, r, s, 1).astype(np.float32) * 2 + 1, axis=2)
    colors = rng.rand(b, r, s, c).astype(np.float32)
    dens = (rng.randn(b, r, s, 1) * 3).astype(np.float32)
    if empty:
        dens[0, 0] = -1e4            # empty ray: weight sum 0, depth nan -> inf -> clamped
    g_rgb = rng.randn(b, r, c).astype(np.float32)
    g_depth = rng.randn(b, r, 1).astype(np.float32)
    g_w = rng.randn(b, r, s - 1, 1).astype(np.float32)
    return colors, dens, depths, g_rgb, g_depth, g_w


def torch_march_grads(colors, dens, depths, g_rgb, g_depth, g_w, white_back, device='cpu'):
    from pix2pix3d_b200.training.volumetric_rendering.ray_marcher import _march_torch
    c = torch.from_numpy(colors).to(device).requires_grad_(True)
    s = torch.from_numpy(dens).to(device).requires_grad_(True)
    rgb, depth, w = _march_torch(c, s, torch.from_numpy(depths).to(device), white_back)
    loss = (rgb * torch.from_numpy(g_rgb).to(device)).sum() + (w * torch.from_numpy(g_w).to(device)).sum()
    if g_depth is not None:
        loss = loss + (depth * torch.from_numpy(g_depth).to(device)).sum()
    loss.backward()
    return c.grad.cpu().numpy(), s.grad.cpu().numpy()


@pytest.mark.parametrize('white_back', [False, True])
@pytest.mark.parametrize('with_depth', [False, True])
def test_oracle_ray_march_backward_matches_autograd(white_back, with_depth):
    colors, dens, depths, g_rgb, g_depth, g_w = march_case(0, empty=not with_depth)   # 0/0 in autograd's depth term is NaN
    gd = g_depth if with_depth else None
    gc, gs = O.renderer.ray_march_backward(colors, dens, depths, g_rgb, gd, g_w, white_back)
    rc, rs = torch_march_grads(colors, dens, depths, g_rgb, gd, g_w, white_back)
    assert rel_err(gc, rc) < 1e-5
    assert rel_err(gs, rs) < 2e-5


def plane_case(seed, b=2, h=12, w=10, m=300):
    rng = np.random.RandomState(seed)
    planes = rng.randn(b, 3, 32, h, w).astype(np.float32)
    coords = ((rng.rand(b, m, 3) - 0.5) * 1.4).astype(np.float32)     # some points outside the box
    g = rng.randn(b, 3, m, 32).astype(np.float32)
    return planes, coords, g


def test_oracle_sample_from_planes_backward_matches_autograd():
    from pix2pix3d_b200.training.volumetric_rendering.renderer import generate_planes, sample_from_planes
    planes, coords, g = plane_case(1)
    p = torch.from_numpy(planes).requires_grad_(True)
    out = sample_from_planes(generate_planes(), p, torch.from_numpy(coords), box_warp=1.0)
    out.backward(torch.from_numpy(g))
    got = O.renderer.sample_from_planes_backward(g, planes.shape, coords, 1.0)
    assert rel_err(got, p.grad.numpy()) < 1e-5


# ---------------------------------------------------------------------------------------------
# GPU: the backward kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('white_back', [False, True])
@pytest.mark.parametrize('shape', [(2, 9, 13, 5), (1, 33, 96, 64), (2, 5, 2, 3), (1, 3, 256, 40)])
def test_ray_march_backward_kernel(shape, white_back):
    from pix2pix3d_b200.training.volumetric_rendering.ray_marcher import MipRayMarcher2
    b, r, s, c = shape
    colors, dens, depths, g_rgb, g_depth, g_w = march_case(2, b, r, s, c, empty=True)
    opts = {'clamp_mode': 'softplus', 'white_back': white_back}
    ct = torch.from_numpy(colors).cuda().requires_grad_(True)
    st = torch.from_numpy(dens).cuda().requires_grad_(True)
    rgb, depth, w = MipRayMarcher2()(ct, st, torch.from_numpy(depths).cuda(), opts)
    # forward values are the forward kernel's
    orgb, odepth, ow = O.renderer.ray_march(colors, dens, depths, white_back)
    assert rel_err(rgb.detach().cpu().numpy(), orgb) < 2e-5
    # 1) rgb + weights only (the training configuration: the depth image is not in the loss)
    loss = (rgb * torch.from_numpy(g_rgb).cuda()).sum() + (w * torch.from_numpy(g_w).cuda()).sum()
    loss.backward(retain_graph=True)
    gc, gs = O.renderer.ray_march_backward(colors, dens, depths, g_rgb, None, g_w, white_back)
    assert rel_err(ct.grad.cpu().numpy(), gc) < 2e-5
    assert rel_err(st.grad.cpu().numpy(), gs) < 1e-4
    rc, rs = torch_march_grads(colors, dens, depths, g_rgb, None, g_w, white_back, device='cuda')
    assert rel_err(ct.grad.cpu().numpy(), rc) < 2e-5
    assert rel_err(st.grad.cpu().numpy(), rs) < 1e-4
    # 2) with a depth gradient (zero on the empty ray, where autograd itself yields NaN)
    ct.grad = None; st.grad = None
    g_depth[0, 0] = 0
    (depth * torch.from_numpy(g_depth).cuda()).sum().backward()
    zc, zs = O.renderer.ray_march_backward(colors, dens, depths, np.zeros_like(g_rgb), g_depth, None, white_back)
    # with a single interval the composite depth is the interval's midpoint and the gradient vanishes identically: compare on an
    # absolute scale there
    assert np.abs(st.grad.cpu().numpy() - zs).max() <= 1e-4 * max(np.abs(zs).max(), 1e-2)
    assert float(ct.grad.abs().max()) == 0.0 and float(np.abs(zc).max()) == 0.0


@pytest.mark.gpu
def test_sample_from_planes_backward_kernel():
    from pix2pix3d_b200.training.volumetric_rendering.renderer import generate_planes, sample_from_planes
    for seed, (b, h, w, m) in enumerate([(2, 12, 10, 300), (1, 64, 64, 5000), (1, 256, 256, 20000)]):
        planes, coords, g = plane_case(seed, b, h, w, m)
        p = torch.from_numpy(planes).cuda().requires_grad_(True)
        out = sample_from_planes(generate_planes().cuda(), p, torch.from_numpy(coords).cuda(), box_warp=1.0)
        assert rel_err(out.detach().cpu().numpy(), O.renderer.sample_from_planes(planes, coords, 1.0)) < 1e-6
        out.backward(torch.from_numpy(g).cuda())
        want = O.renderer.sample_from_planes_backward(g, planes.shape, coords, 1.0)
        assert rel_err(p.grad.cpu().numpy(), want) < 1e-5          # fp32 atomics: order-dependent rounding only


@pytest.mark.gpu
def test_staged_renderer_gradients_match_aten_autograd():
    """ImportanceRenderer.forward under autograd (the training path): gradients w.r.t. planes and decoder weights through the
    kernel-backed stages against the same pipeline on ATen ops (grid_sample + torch ray marcher), same random draws."""
    from pix2pix3d_b200.training.triplane import OSGDecoder
    from pix2pix3d_b200.training.volumetric_rendering import ray_marcher as rm
    from pix2pix3d_b200.training.volumetric_rendering import renderer as rr
    dev = torch.device('cuda')
    torch.manual_seed(0)
    dec = OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32}).to(dev)
    planes0 = torch.randn(2, 3, 32, 32, 32, device=dev)
    o = torch.zeros(2, 50, 3, device=dev); o[..., 2] = 2.7
    d = torch.nn.functional.normalize(torch.randn(2, 50, 3, device=dev) * 0.15 - torch.tensor([0, 0, 1.0], device=dev), dim=-1)
    opts = {'ray_start': 2.25, 'ray_end': 3.3, 'depth_resolution': 12, 'depth_resolution_importance': 12, 'box_warp': 1.0,
            'disparity_space_sampling': False, 'clamp_mode': 'softplus', 'white_back': False, 'density_noise': 0}
    g_feat = torch.randn(2, 50, 32, device=dev)
    g_ws = torch.randn(2, 50, 1, device=dev)

    def run(use_kernels):
        renderer = rr.ImportanceRenderer()
        planes = planes0.clone().requires_grad_(True)
        dec.zero_grad()
        torch.manual_seed(1)
        if use_kernels:
            feat, depth, wsum = renderer(planes, dec, o, d, opts)
        else:
            saved = (rr._SamplePlanes.apply, rm._RayMarch.apply)
            rr._SamplePlanes.apply = staticmethod(lambda pf, co, bw: _aten_sample(rr, pf, co, bw))
            rm._RayMarch.apply = staticmethod(lambda c, s, z, wb: rm._march_torch(c, s, z, wb))
            try:
                feat, depth, wsum = renderer(planes, dec, o, d, opts)
            finally:
                rr._SamplePlanes.apply, rm._RayMarch.apply = saved
        ((feat * g_feat).sum() + (wsum * g_ws).sum()).backward()
        return feat.detach(), planes.grad.clone(), [p.grad.clone() for p in dec.parameters()]

    f1, gp1, gd1 = run(True)
    f2, gp2, gd2 = run(False)
    assert rel_err(f1.cpu().numpy(), f2.cpu().numpy()) < 1e-4
    assert rel_err(gp1.cpu().numpy(), gp2.cpu().numpy()) < 2e-3
    for a, b in zip(gd1, gd2):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 2e-3


def _aten_sample(rr, plane_features, coordinates, box_warp):
    n, n_planes, c, h, w = plane_features.shape
    m = coordinates.shape[1]
    feats = plane_features.reshape(n * n_planes, c, h, w)
    grid = rr.project_onto_planes(rr.generate_planes().to(coordinates.device), (2 / box_warp) * coordinates).unsqueeze(1)
    out = torch.nn.functional.grid_sample(feats, grid.float(), mode='bilinear', padding_mode='zeros', align_corners=False)
    return out.permute(0, 3, 2, 1).reshape(n, n_planes, m, c)


# ---------------------------------------------------------------------------------------------
# OSG decoder: oracle backward (hand-derived) vs autograd of the decoder modules (CPU), CUDA kernels vs the oracle (GPU)
# ---------------------------------------------------------------------------------------------
def _decoder_case(kind, seed=5, n=2, m=37, lr=1.0):
    from pix2pix3d_b200.training import triplane, triplane_cond as tc
    base = {'decoder_lr_mul': lr, 'decoder_output_dim': 32}
    torch.manual_seed(seed)
    if kind == 'osg':
        mod, okind, sig = triplane.OSGDecoder(32, dict(base)), 'OSGDecoder', None
    elif kind in ('semantic_raw', 'semantic_sigmoid'):
        sig = kind.endswith('sigmoid')
        mod, okind = tc.OSGDecoder_semantic(32, dict(base, sigmoid=sig)), 'OSGDecoder_semantic'
    else:
        sig = kind == 'late1'
        mod, okind = tc.OSGDecoder_semantic_lateSeparate(32, dict(base, sigmoid=sig, semantic_channels=1 if sig else 6)), \
            'OSGDecoder_semantic_lateSeparate'
    for p in mod.parameters():
        p.data.normal_(0, 0.7)
    nets = [mod.net] + ([mod.net_semantic] if okind.endswith('lateSeparate') else [])
    dec = dict(kind=okind, lr_mul=lr, sigmoid=sig,
               nets=[dict(w1=t[0].weight.detach().numpy(), b1=t[0].bias.detach().numpy(), w2=t[2].weight.detach().numpy(),
                          b2=t[2].bias.detach().numpy()) for t in nets])
    feats = torch.randn(n, 3, m, 32) * 2
    co = 32 * len(nets)
    return mod, nets, dec, feats, torch.randn(n, m, co), torch.randn(n, m, 1)


DEC_KINDS = ['osg', 'semantic_raw', 'semantic_sigmoid', 'late6', 'late1']


@pytest.mark.parametrize('kind', DEC_KINDS)
@pytest.mark.parametrize('lr', [1.0, 0.5])
def test_oracle_decoder_backward_matches_autograd(kind, lr):
    mod, nets, dec, feats, g_rgb, g_sigma = _decoder_case(kind, lr=lr)
    mod = mod.double()
    f = feats.double().requires_grad_(True)
    out = mod(f, None)
    params = [p for t in nets for p in (t[0].weight, t[0].bias, t[2].weight, t[2].bias)]
    want = torch.autograd.grad((out['rgb'] * g_rgb.double()).sum() + (out['sigma'] * g_sigma.double()).sum(), [f] + params)
    rgb, sigma = O.renderer.decoder_forward(dec, feats.numpy())
    assert rel_err(rgb, out['rgb'].detach().numpy()) < 1e-5 and rel_err(sigma, out['sigma'].detach().numpy()) < 1e-5
    gf, grads = O.renderer.decoder_backward(dec, feats.numpy(), g_rgb.numpy(), g_sigma.numpy())
    assert rel_err(gf, want[0].numpy()) < 1e-10
    got = [g[q] for g in grads for q in ('w1', 'b1', 'w2', 'b2')]
    for a, b in zip(got, want[1:]):
        assert rel_err(a, b.numpy()) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize('kind', DEC_KINDS)
def test_decoder_kernels_match_the_oracle(kind):
    """p3d_decoder_mlp_fwd / _bwd through the decoder modules on CUDA against the oracle's forward and backward."""
    mod, nets, dec, feats, g_rgb, g_sigma = _decoder_case(kind, n=2, m=1000)
    mod = mod.cuda()
    f = feats.cuda().requires_grad_(True)
    out = mod(f, None)
    params = [p for t in nets for p in (t[0].weight, t[0].bias, t[2].weight, t[2].bias)]
    got = torch.autograd.grad((out['rgb'] * g_rgb.cuda()).sum() + (out['sigma'] * g_sigma.cuda()).sum(), [f] + params)
    rgb, sigma = O.renderer.decoder_forward(dec, feats.numpy())
    assert rel_err(out['rgb'].detach().cpu().numpy(), rgb) < 1e-5 and rel_err(out['sigma'].detach().cpu().numpy(), sigma) < 1e-5
    gf, grads = O.renderer.decoder_backward(dec, feats.numpy(), g_rgb.numpy(), g_sigma.numpy())
    assert rel_err(got[0].cpu().numpy(), gf) < 2e-5
    want = [g[q] for g in grads for q in ('w1', 'b1', 'w2', 'b2')]
    for a, b in zip(got[1:], want):
        assert rel_err(a.cpu().numpy(), b) < 2e-5
