"""CUDA renderer kernels (through the C-ABI, via pix2pix3d_b200.native) against the oracle and the reference
fixtures. Tolerance: north_star asks 1e-3 relative fp32 on outputs and bit-exact sampling bookkeeping."""
import numpy as np
import pytest
import torch

import p3d_oracle as O
from conftest import load_golden, rel_err
from test_oracle_golden import RENDER_CASES, TC_RENDER_CASES, oracle_decoder, render_opts

pytestmark = pytest.mark.gpu
TOL = 1e-3          # north_star tolerance
TIGHT = 2e-5        # what the fp32 kernels actually achieve against the oracle


class _Dec(torch.nn.Module):
    """Stand-in with the attribute layout native.describe_decoder expects, built from fixture weights."""


def torch_decoder(g, device):
    from pix2pix3d_b200.training.triplane import OSGDecoder
    from pix2pix3d_b200.training.triplane_cond import OSGDecoder_semantic_lateSeparate
    kind = str(g['decoder'])
    if kind == 'osg':
        dec = OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32})
    else:
        cs = 6 if kind == 'late6' else 1
        dec = OSGDecoder_semantic_lateSeparate(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32, 'sigmoid': cs == 1,
                                                    'semantic_channels': cs})
    dec.load_state_dict({k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('dec.')})
    return dec.to(device).requires_grad_(False)


def run_fused(g, debug=True, impl='simt'):
    from pix2pix3d_b200 import native
    dev = torch.device('cuda')
    opts = render_opts(g)
    b, m = g['ray_origins'].shape[:2]
    dc = O.renderer.sample_stratified(b, m, opts['ray_start'], opts['ray_end'], opts['depth_resolution'], g['jitter'])
    planes = torch.from_numpy(g['planes']).to(dev)
    dec = native.pack_decoder(torch_decoder(g, dev))
    u = torch.from_numpy(g['u']).to(dev) if int(g['Sf']) > 0 else None
    res = native.render_fwd(native.planes_to_channels_last(planes), dec, torch.from_numpy(g['ray_origins']).to(dev),
                            torch.from_numpy(g['ray_dirs']).to(dev), torch.from_numpy(dc).to(dev), u, opts['box_warp'],
                            white_back=opts['white_back'], debug=debug, impl=impl)
    torch.cuda.synchronize()
    return res, dc, opts


IMPL_CASES = ([(c, 'simt') for c in RENDER_CASES] + [(c, 'tc') for c in TC_RENDER_CASES]
              + [(c, 'tc_pairs') for c in TC_RENDER_CASES])      # every tc fixture has <= 64 samples per pass


@pytest.mark.parametrize('case,impl', IMPL_CASES)
def test_fused_render_matches_reference_and_oracle(case, impl):
    g = load_golden('renderer_' + case)
    (feat, depth, wsum, dbg), dc, opts = run_fused(g, impl=impl)
    assert rel_err(feat.cpu().numpy(), g['feat']) < TOL
    assert rel_err(depth.cpu().numpy(), g['depth']) < TOL
    assert rel_err(wsum.cpu().numpy(), g['wsum']) < TOL
    u = g['u'] if int(g['Sf']) > 0 else None
    of, od, ow, odbg = O.renderer.importance_renderer(g['planes'], oracle_decoder(g), g['ray_origins'], g['ray_dirs'], dc, u, opts,
                                                      return_debug=True)
    assert rel_err(feat.cpu().numpy(), of) < TIGHT
    assert rel_err(depth.cpu().numpy(), od) < TIGHT
    assert rel_err(wsum.cpu().numpy()[..., 0], ow[..., 0] if ow.ndim == 3 else ow) < 1e-4   # tiny sums of 1-exp(-x)
    assert rel_err(dbg['weights_final'].cpu().numpy(), odbg['weights_final']) < 2e-4


@pytest.mark.parametrize('case,impl', [ci for ci in IMPL_CASES if ci[0] not in ('coarse_only', 'coarse8')])
def test_sampling_bookkeeping_is_bit_exact(case, impl):
    """Feed the oracle the kernel's OWN coarse weights: searchsorted indices, fine depths and the sort permutation
    must then be identical bit for bit (integer bookkeeping of renderer.py:240-252 and :162)."""
    g = load_golden('renderer_' + case)
    (feat, depth, wsum, dbg), dc, opts = run_fused(g, impl=impl)
    b, m = g['ray_origins'].shape[:2]
    sc, sf = opts['depth_resolution'], opts['depth_resolution_importance']
    wc = dbg['weights_coarse'].cpu().numpy()
    fine, odbg = O.renderer.sample_importance(dc.reshape(b * m, sc), wc.reshape(b * m, sc - 1), g['u'], return_debug=True)
    assert np.array_equal(dbg['inds'].cpu().numpy().reshape(b * m, sf), odbg['inds'])
    kf = dbg['depths_fine'].cpu().numpy().reshape(b * m, sf)
    assert np.array_equal(kf.view(np.uint32), fine.view(np.uint32))
    alld = np.concatenate([dc.reshape(b, m, sc), kf.reshape(b, m, sf)], -1)
    perm = np.argsort(alld, axis=-1, kind='stable')
    assert np.array_equal(dbg['perm'].cpu().numpy(), perm)
    # the standalone entry point agrees with the fused kernel
    from pix2pix3d_b200 import native
    s2, i2 = native.sample_importance(torch.from_numpy(dc.reshape(b * m, sc)).cuda(), dbg['weights_coarse'].reshape(b * m, sc - 1),
                                      torch.from_numpy(g['u']).cuda(), return_inds=True)
    assert np.array_equal(s2.cpu().numpy().view(np.uint32), kf.view(np.uint32))
    assert np.array_equal(i2.cpu().numpy(), odbg['inds'])


def test_ray_sampler_matches_oracle():
    from pix2pix3d_b200 import native
    g = load_golden('renderer_seg')
    for res in (12, 64, 128):
        o, d = native.ray_sampler(torch.from_numpy(g['cam2world']).cuda(), torch.from_numpy(g['intrinsics']).cuda(), res)
        oo, dd = O.renderer.ray_sampler(g['cam2world'], g['intrinsics'], res)
        assert np.array_equal(o.cpu().numpy(), oo)
        assert np.abs(d.cpu().numpy() - dd).max() < 3e-7
    o, d = native.ray_sampler(torch.from_numpy(g['cam2world']).cuda(), torch.from_numpy(g['intrinsics']).cuda(), int(g['nrr']))
    assert np.abs(d.cpu().numpy() - g['ray_dirs']).max() < 3e-7


@pytest.mark.parametrize('impl', ['simt', 'tc'])
def test_run_model_and_sample_from_planes(impl):
    from pix2pix3d_b200 import native
    g = load_golden('renderer_seg')
    dev = torch.device('cuda')
    rng = np.random.RandomState(0)
    coords = (rng.rand(2, 777, 3).astype(np.float32) - 0.5) * 1.3      # some points fall outside the box
    planes = torch.from_numpy(g['planes']).to(dev)
    pcl = native.planes_to_channels_last(planes)
    assert torch.equal(pcl, planes.permute(0, 1, 3, 4, 2).contiguous())
    feats = native.sample_from_planes(pcl, torch.from_numpy(coords).to(dev), 1.0)
    ofe = O.renderer.sample_from_planes(g['planes'], coords, 1.0)
    assert rel_err(feats.cpu().numpy(), ofe) < 1e-6
    dec = native.pack_decoder(torch_decoder(g, dev))
    rgb, sigma = native.run_model(pcl, dec, torch.from_numpy(coords).to(dev), 1.0, impl=impl)
    orgb, osig = O.renderer.run_model(g['planes'], oracle_decoder(g), coords, 1.0)
    assert rel_err(rgb.cpu().numpy(), orgb) < TIGHT
    assert rel_err(sigma.cpu().numpy(), osig) < TIGHT


@pytest.mark.parametrize('case', ['seg', 'car', 'rgb_only', 'far_outside'])
@pytest.mark.parametrize('m', [1, 127, 128, 129, 5000])
def test_run_model_tc_decoders_and_ragged_sizes(case, m):
    """Tensor-core point query (p3d_run_model_tc) against the oracle for every decoder family, with point counts around the
    128-row tile size and points far outside the box (zero padding)."""
    from pix2pix3d_b200 import native
    g = load_golden('renderer_' + case)
    dev = torch.device('cuda')
    b = g['planes'].shape[0]
    rng = np.random.RandomState(m)
    coords = (rng.rand(b, m, 3).astype(np.float32) - 0.5) * 2.6
    if m > 4:
        coords[0, 3] = 1e6
        coords[-1, m // 2, 1] = -3e4
    box = 1.0 if case != 'car' else 1.6
    pcl = native.planes_to_channels_last(torch.from_numpy(g['planes']).to(dev))
    dec = native.pack_decoder(torch_decoder(g, dev))
    rgb, sigma = native.run_model(pcl, dec, torch.from_numpy(coords).to(dev), box, impl='tc')
    orgb, osig = O.renderer.run_model(g['planes'], oracle_decoder(g), coords, box)
    assert rgb.shape == (b, m, orgb.shape[-1]) and sigma.shape == (b, m, 1)
    assert rel_err(rgb.cpu().numpy(), orgb) < TIGHT
    assert rel_err(sigma.cpu().numpy(), osig) < TIGHT
    # and against the CUDA-core kernel: same gather, fp32 decoder
    rgb2, sigma2 = native.run_model(pcl, dec, torch.from_numpy(coords).to(dev), box, impl='simt')
    assert rel_err(rgb.cpu().numpy(), rgb2.cpu().numpy()) < TIGHT
    assert rel_err(sigma.cpu().numpy(), sigma2.cpu().numpy()) < TIGHT


def test_run_model_tc_dense_grid_and_strided_planes():
    """The extract_mesh pattern (applications/extract_mesh.py:60-81 in the reference): a 64^3 block of grid points per call
    against full-size 256^2 planes, here read in place from the backbone's NHWC [B,256,256,96] layout. The tensor-core
    query must agree with the CUDA-core kernel on dense planes; per-point results do not depend on the batch they ran in."""
    from pix2pix3d_b200 import native
    g = load_golden('renderer_seg')
    dev = torch.device('cuda')
    gen = torch.Generator(device='cpu').manual_seed(5)
    img = torch.randn(2, 256, 256, 96, generator=gen).to(dev)           # backbone output, channels last
    view = img.view(2, 256, 256, 3, 32).permute(0, 3, 1, 2, 4)             # [B,3,H,W,32] strided view
    assert not view.is_contiguous()
    dec = native.pack_decoder(torch_decoder(g, dev))
    n = 64
    ax = torch.linspace(-0.55, 0.55, n)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing='ij'), -1).reshape(1, -1, 3).repeat(2, 1, 1).to(dev)
    rgb, sigma = native.run_model(view, dec, grid, 1.0, impl='tc')
    rgb2, sigma2 = native.run_model(view.contiguous(), dec, grid, 1.0, impl='simt')
    assert rel_err(rgb.cpu().numpy(), rgb2.cpu().numpy()) < TIGHT
    assert rel_err(sigma.cpu().numpy(), sigma2.cpu().numpy()) < TIGHT
    part, sp = native.run_model(view[1:], dec, grid[1:, 1000:1777], 1.0, impl='tc')
    assert torch.equal(part, rgb[1:, 1000:1777]) and torch.equal(sp, sigma[1:, 1000:1777])
    # densities only (what extract_mesh keeps): same sigma bit for bit, no colour buffer
    none, so = native.run_model(view, dec, grid, 1.0, impl='tc', sigma_only=True)
    assert none is None and torch.equal(so, sigma)


def test_ray_march_standalone_matches_oracle():
    from pix2pix3d_b200 import native
    rng = np.random.RandomState(1)
    b, r, s, c = 2, 37, 23, 7
    depths = np.sort(rng.rand(b, r, s, 1).astype(np.float32) * 2 + 1, axis=2)
    colors = rng.rand(b, r, s, c).astype(np.float32)
    dens = rng.randn(b, r, s, 1).astype(np.float32) * 3
    dens[0, 0] = -1e4                                     # empty ray: weight_total == 0 -> nan -> clamp to max depth
    for wb in (False, True):
        rgb, depth, w = native.ray_march(torch.from_numpy(colors).cuda(), torch.from_numpy(dens).cuda(),
                                         torch.from_numpy(depths).cuda(), wb)
        orgb, odepth, ow = O.renderer.ray_march(colors, dens, depths, wb)
        assert rel_err(rgb.cpu().numpy(), orgb) < TIGHT
        assert rel_err(w.cpu().numpy(), ow) < 1e-4
        assert rel_err(depth.cpu().numpy(), odepth) < TIGHT
        assert depth[0, 0, 0].item() == depths.max()


@pytest.mark.parametrize('impl', ['simt', 'tc', 'tc_pairs'])
def test_full_size_properties_config2(impl):
    """BASELINE config 2 sizes (B=4, 128^2 rays, 48+48 samples): size-independent properties."""
    from pix2pix3d_b200 import native
    from pix2pix3d_b200.training.triplane_cond import OSGDecoder_semantic_lateSeparate
    dev = torch.device('cuda')
    torch.manual_seed(0)
    B, H, nrr, Sc, Sf = 4, 256, 128, 48, 48
    planes = torch.randn(B, 3, 32, H, H, device=dev)
    dec_m = OSGDecoder_semantic_lateSeparate(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32, 'sigmoid': False,
                                                  'semantic_channels': 6}).to(dev).requires_grad_(False)
    g = load_golden('renderer_seg')
    c2w = torch.from_numpy(g['cam2world'][:1]).to(dev).repeat(B, 1, 1)
    K = torch.from_numpy(g['intrinsics'][:1]).to(dev).repeat(B, 1, 1)
    o, d = native.ray_sampler(c2w, K, nrr)
    R = nrr * nrr
    base = torch.linspace(2.25, 3.3, Sc, device=dev).reshape(1, 1, Sc)
    dc = base + torch.rand(B, R, Sc, device=dev) * ((3.3 - 2.25) / (Sc - 1))
    u = torch.rand(B * R, Sf, device=dev)
    dec = native.pack_decoder(dec_m)
    pcl = native.planes_to_channels_last(planes)
    feat, depth, wsum, dbg = native.render_fwd(pcl, dec, o, d, dc, u, 1.0, debug=True, impl=impl)
    feat2, depth2, wsum2 = native.render_fwd(pcl, dec, o, d, dc, u, 1.0, impl=impl)
    torch.cuda.synchronize()
    assert torch.equal(feat, feat2) and torch.equal(depth, depth2) and torch.equal(wsum, wsum2)   # deterministic
    assert torch.isfinite(feat).all() and torch.isfinite(depth).all()
    assert (wsum >= 0).all() and (wsum <= 1 + 1e-5).all()
    assert torch.allclose(dbg['weights_final'].sum(-1, keepdim=True), wsum, atol=1e-5)
    perm = dbg['perm'].long()
    assert torch.equal(perm.sort(-1)[0], torch.arange(Sc + Sf, device=dev).expand_as(perm))        # a permutation
    alld = torch.cat([dc, dbg['depths_fine']], -1)
    sd = torch.gather(alld, -1, perm)
    assert (sd[..., 1:] >= sd[..., :-1]).all()                                                    # sorted
    assert depth.min() >= alld.min() and depth.max() <= alld.max()
    assert (dbg['inds'] >= 1).all() and (dbg['inds'] <= Sc - 2).all()
    # image 2 rendered alone equals its slice of the batch (rays are independent; depth only via the global clamp)
    f1, d1, w1 = native.render_fwd(pcl[2:3].contiguous(), dec, o[2:3].contiguous(), d[2:3].contiguous(), dc[2:3].contiguous(),
                                   u[2 * R:3 * R].contiguous(), 1.0, impl=impl)
    assert torch.equal(f1[0], feat[2]) and torch.equal(w1[0], wsum[2])
    # the backbone's NHWC [B,H,W,96] output read in place as a strided [B,3,H,W,32] view (plane_strides of the C-ABI)
    inter = pcl.permute(0, 2, 3, 1, 4).contiguous()
    view = inter.permute(0, 3, 1, 2, 4)
    assert not view.is_contiguous()
    fs, dsx, wsx = native.render_fwd(view, dec, o, d, dc, u, 1.0, impl=impl)
    assert torch.equal(fs, feat) and torch.equal(dsx, depth) and torch.equal(wsx, wsum)
    # sampled check of full-size output against the oracle on 64 rays
    idx = torch.randperm(R, device=dev)[:64]
    pl = planes[1:2].cpu().numpy()
    from test_oracle_golden import oracle_decoder as _od
    sdict = {('dec.' + k): v.cpu().numpy() for k, v in dec_m.state_dict().items()}
    sdict['decoder'] = np.array('late6')
    of, odp, ow = O.renderer.importance_renderer(pl, _od(sdict), o[1:2, idx].cpu().numpy(), d[1:2, idx].cpu().numpy(),
                                                 dc[1:2, idx].cpu().numpy()[..., None], u.reshape(B, R, Sf)[1, idx].cpu().numpy(),
                                                 dict(box_warp=1.0))
    assert rel_err(feat[1:2, idx].cpu().numpy(), of) < TIGHT
    assert rel_err(wsum[1:2, idx].cpu().numpy(), ow) < TIGHT


def test_importance_renderer_module_dispatches_to_fused_kernel():
    """The reference-facing call (ImportanceRenderer.forward with NCHW planes, decoder module, options dict)."""
    from pix2pix3d_b200 import _lib
    from pix2pix3d_b200.training.volumetric_rendering.renderer import ImportanceRenderer
    g = load_golden('renderer_car')
    dev = torch.device('cuda')
    opts = dict(render_opts(g), clamp_mode='softplus', disparity_space_sampling=False)
    it = iter([torch.from_numpy(g['jitter']).to(dev), torch.from_numpy(g['u']).to(dev)])
    o_like, o_rand = torch.rand_like, torch.rand
    torch.rand_like = lambda x, *a, **k: next(it)
    torch.rand = lambda *a, **k: next(it)
    before = _lib.launch_count
    try:
        with torch.no_grad():
            feat, depth, wsum = ImportanceRenderer()(torch.from_numpy(g['planes']).to(dev), torch_decoder(g, dev),
                                                     torch.from_numpy(g['ray_origins']).to(dev),
                                                     torch.from_numpy(g['ray_dirs']).to(dev), opts)
    finally:
        torch.rand_like, torch.rand = o_like, o_rand
    assert _lib.launch_count - before == 4          # pack_decoder (2 images), planes transpose, fused render
    assert rel_err(feat.cpu().numpy(), g['feat']) < TOL
    assert rel_err(depth.cpu().numpy(), g['depth']) < TOL
    assert rel_err(wsum.cpu().numpy(), g['wsum']) < TOL


@pytest.mark.parametrize('impl', ['simt', 'tc', 'tc_pairs'])
def test_unsorted_and_tied_coarse_depths_still_merge_like_a_stable_sort(impl):
    """The C-ABI accepts any depths_coarse. The tensor-core kernel takes a shortcut when a ray's coarse depths are
    non-decreasing (what sample_stratified produces); rays that are not, and rays with exact ties, must still produce the
    stable sort permutation of torch.sort / np.argsort(kind='stable') (renderer.py:162)."""
    from pix2pix3d_b200 import native
    g = load_golden('renderer_seg16')
    dev = torch.device('cuda')
    opts = render_opts(g)
    b, m = g['ray_origins'].shape[:2]
    sc, sf = int(g['Sc']), int(g['Sf'])
    dc = O.renderer.sample_stratified(b, m, opts['ray_start'], opts['ray_end'], sc, g['jitter']).reshape(b, m, sc).copy()
    dc[:, 0::3, [3, 4]] = dc[:, 0::3, [4, 3]]            # every third ray: one inversion
    dc[:, 1::3, 7] = dc[:, 1::3, 6]                      # every third ray: an exact tie (still non-decreasing)
    planes = torch.from_numpy(g['planes']).to(dev)
    dec = native.pack_decoder(torch_decoder(g, dev))
    feat, depth, wsum, dbg = native.render_fwd(native.planes_to_channels_last(planes), dec, torch.from_numpy(g['ray_origins']).to(dev),
                                               torch.from_numpy(g['ray_dirs']).to(dev), torch.from_numpy(dc[..., None]).to(dev),
                                               torch.from_numpy(g['u']).to(dev), opts['box_warp'], debug=True, impl=impl)
    torch.cuda.synchronize()
    kf = dbg['depths_fine'].cpu().numpy().reshape(b, m, sf)
    alld = np.concatenate([dc, kf], -1)
    perm = np.argsort(alld, axis=-1, kind='stable')
    assert np.array_equal(dbg['perm'].cpu().numpy(), perm)
    assert torch.isfinite(feat).all() and torch.isfinite(depth).all()


# ---------------------------------------------------------------------------------------------------------------
# Bookkeeping scored against the REFERENCE's own torch.searchsorted / torch.sort results stored in the fixtures
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case,impl', [ci for ci in IMPL_CASES if ci[0] not in ('coarse_only', 'coarse8')])
def test_bookkeeping_against_reference_fixture(case, impl):
    g = load_golden('renderer_' + case)
    (feat, depth, wsum, dbg), dc, opts = run_fused(g, impl=impl)
    perm = dbg['perm'].cpu().numpy().reshape(g['perm'].shape)
    dfine = dbg['depths_fine'].cpu().numpy().reshape(g['depths_fine'].shape)
    perm_rate = float((perm == g['perm']).mean())
    print(f'BOOKKEEPING-VS-REFERENCE {case}/{impl}: perm exact {perm_rate:.5f}, depths_fine rel err {rel_err(dfine, g["depths_fine"]):.2e}')
    assert rel_err(dfine, g['depths_fine']) < 1e-5
    assert perm_rate >= 0.999


# ---------------------------------------------------------------------------------------------------------------
# sample_stratified inside the kernel (depth_mode 1 / 2), ray limits, plane-set indirection
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('case,impl', [('seg', 'simt'), ('seg48', 'tc'), ('car64', 'tc_pairs')])
def test_in_kernel_stratified_depths_are_bit_identical(case, impl):
    """depth_mode 1: the kernel forms linspace[k] + jitter * delta itself; outputs must equal the run on precomputed depths
    bit for bit, and the depths the reference's sample_stratified produces (torch ops) must be what it formed."""
    from pix2pix3d_b200 import native
    g = load_golden('renderer_' + case)
    dev = torch.device('cuda')
    opts = render_opts(g)
    b, m = g['ray_origins'].shape[:2]
    sc = opts['depth_resolution']
    planes_cl = native.planes_to_channels_last(torch.from_numpy(g['planes']).to(dev))
    dec = native.pack_decoder(torch_decoder(g, dev))
    o, d = torch.from_numpy(g['ray_origins']).to(dev), torch.from_numpy(g['ray_dirs']).to(dev)
    u = torch.from_numpy(g['u']).to(dev)
    jitter = torch.from_numpy(g['jitter']).to(dev)
    # the reference's own op sequence on the device (renderer.py:187-190)
    depths = torch.linspace(opts['ray_start'], opts['ray_end'], sc, device=dev).reshape(1, 1, sc, 1).repeat(b, m, 1, 1)
    depths += jitter * ((opts['ray_end'] - opts['ray_start']) / (sc - 1))
    ref = native.render_fwd(planes_cl, dec, o, d, depths, u, opts['box_warp'], white_back=opts['white_back'], impl=impl, debug=True)
    table = torch.linspace(opts['ray_start'], opts['ray_end'], sc, device=dev)
    got = native.render_fwd(planes_cl, dec, o, d, None, u, opts['box_warp'], white_back=opts['white_back'], impl=impl, debug=True,
                            stratified=dict(jitter=jitter, table=table, delta=(opts['ray_end'] - opts['ray_start']) / (sc - 1)))
    for a, r in zip(got[:3], ref[:3]):
        assert torch.equal(a, r)
    for k in ('perm', 'inds', 'depths_fine', 'weights_coarse'):
        assert torch.equal(got[3][k], ref[3][k]), k


def test_per_ray_limits_in_kernel_and_ray_box_kernel():
    """`ray_start == 'auto'` (renderer.py:91-97): the slab test kernel against the mirror's torch formula evaluated on CPU
    (bit-exact), then depth_mode 2 against the torch evaluation of math_utils.linspace + jitter (bit-exact outputs)."""
    from pix2pix3d_b200 import native
    from pix2pix3d_b200.training.volumetric_rendering import math_utils
    g = load_golden('renderer_far_outside')
    dev = torch.device('cuda')
    opts = render_opts(g)
    o, d = torch.from_numpy(g['ray_origins']), torch.from_numpy(g['ray_dirs'])
    # rays that miss, graze and hit; one direction component exactly zero
    d2 = d.clone(); d2[0, :5, 0] = 0.0; d2[0, 5:9] = torch.tensor([0.0, 0.0, 1.0])
    d2[0, 9:13] = torch.nn.functional.normalize(torch.tensor([[1.0, 0.2, 0.1]]), dim=-1)      # looking sideways: misses the cube
    tn_cpu, tf_cpu = math_utils.get_ray_limits_box(o, d2, box_side_length=1.0)
    tn, tf = native.ray_limits_box(o.to(dev), d2.to(dev), 1.0)
    assert torch.equal(torch.nan_to_num(tn.cpu(), nan=-7), torch.nan_to_num(tn_cpu, nan=-7))
    assert torch.equal(torch.nan_to_num(tf.cpu(), nan=-7), torch.nan_to_num(tf_cpu, nan=-7))
    assert (tn_cpu == -1).any() and (tf_cpu == -2).any() and (tn_cpu > 0).any()

    b, m = o.shape[:2]
    sc = opts['depth_resolution']
    rs, re_ = native.ray_limits_box(o.to(dev), d.to(dev), 1.0)
    valid = re_ > rs
    rs[~valid] = rs[valid].min(); re_[~valid] = rs[valid].max()
    jitter = torch.from_numpy(g['jitter']).to(dev)
    depths = math_utils.linspace(rs, re_, sc).permute(1, 2, 0, 3)
    depths = depths + jitter * ((re_ - rs) / (sc - 1))[..., None]
    planes_cl = native.planes_to_channels_last(torch.from_numpy(g['planes']).to(dev))
    dec = native.pack_decoder(torch_decoder(g, dev))
    u = torch.from_numpy(g['u']).to(dev)
    for impl in ('simt', 'tc'):
        if impl == 'tc' and (sc % 8 or int(g['Sf']) % 8):
            continue
        ref = native.render_fwd(planes_cl, dec, o.to(dev), d.to(dev), depths, u, opts['box_warp'], impl=impl)
        table = torch.arange(sc, dtype=torch.float32, device=dev) / (sc - 1)
        got = native.render_fwd(planes_cl, dec, o.to(dev), d.to(dev), None, u, opts['box_warp'], impl=impl,
                                stratified=dict(jitter=jitter, table=table, ray_start=rs, ray_end=re_))
        for a, r in zip(got, ref):
            assert torch.equal(a, r), impl


@pytest.mark.parametrize('impl', ['simt', 'tc', 'tc_pairs'])
def test_plane_index_shares_one_plane_set_between_views(impl):
    """V camera views of one plane set (generate_video.py:57-69): rays batch V, planes batch 1 + plane_index == the per-view
    renders against a replicated plane tensor."""
    from pix2pix3d_b200 import native
    g = load_golden('renderer_seg16')
    dev = torch.device('cuda')
    opts = render_opts(g)
    planes = torch.from_numpy(g['planes']).to(dev)               # [2,...]: two plane sets
    o, d = torch.from_numpy(g['ray_origins']).to(dev), torch.from_numpy(g['ray_dirs']).to(dev)
    b, m = o.shape[:2]
    dc = torch.from_numpy(O.renderer.sample_stratified(b, m, opts['ray_start'], opts['ray_end'], opts['depth_resolution'], g['jitter'])).to(dev)
    u = torch.from_numpy(g['u']).to(dev)
    dec = native.pack_decoder(torch_decoder(g, dev))
    # views: image 0 and 1 both look at plane set 1; reference = planes[[1, 1]]
    ref = native.render_fwd(native.planes_to_channels_last(planes[[1, 1]].contiguous()), dec, o, d, dc, u, opts['box_warp'], impl=impl)
    got = native.render_fwd(native.planes_to_channels_last(planes[1:2].contiguous()), dec, o, d, dc, u, opts['box_warp'], impl=impl,
                            plane_index=torch.zeros(2, dtype=torch.int32))
    for a, r in zip(got, ref):
        assert torch.equal(a, r)
    mixed = native.render_fwd(native.planes_to_channels_last(planes), dec, o, d, dc, u, opts['box_warp'], impl=impl,
                              plane_index=torch.tensor([1, 0], dtype=torch.int32))
    swap = native.render_fwd(native.planes_to_channels_last(planes[[1, 0]].contiguous()), dec, o, d, dc, u, opts['box_warp'], impl=impl)
    assert torch.equal(mixed[0], swap[0])


@pytest.mark.parametrize('kind', ['osg', 'semantic_raw', 'semantic_sigmoid', 'entangle_6', 'entangle_19', 'entangle_sigmoid', 'late_19'])
@pytest.mark.parametrize('impl', ['tc', 'simt'])
def test_fused_renderer_decoder_kinds_match_the_staged_modules(kind, impl, monkeypatch):
    """Every decoder layout native.describe_decoder maps onto the fused kernels (one or two nets, per-channel sigmoid masks:
    triplane.py:112-135, triplane_cond.py:859-970) against ImportanceRenderer's staged path, which evaluates the decoder MODULE
    (mirror of the reference's forward) between the stage kernels, on the same rays and the same random draws."""
    from pix2pix3d_b200 import native
    from pix2pix3d_b200.training import triplane, triplane_cond as tc
    from pix2pix3d_b200.training.volumetric_rendering.renderer import ImportanceRenderer
    dev = torch.device('cuda')
    base = {'decoder_lr_mul': 1.0, 'decoder_output_dim': 32}
    make = {
        'osg': lambda: triplane.OSGDecoder(32, dict(base)),
        'semantic_raw': lambda: tc.OSGDecoder_semantic(32, dict(base, sigmoid=False)),
        'semantic_sigmoid': lambda: tc.OSGDecoder_semantic(32, dict(base, sigmoid=True)),
        'entangle_6': lambda: tc.OSGDecoder_semantic_entangle(32, dict(base, sigmoid=False, semantic_channels=6)),
        'entangle_19': lambda: tc.OSGDecoder_semantic_entangle(32, dict(base, sigmoid=False, semantic_channels=19)),
        'entangle_sigmoid': lambda: tc.OSGDecoder_semantic_entangle(32, dict(base, sigmoid=True, semantic_channels=1)),
        'late_19': lambda: tc.OSGDecoder_semantic_lateSeparate(32, dict(base, sigmoid=False, semantic_channels=19)),
    }[kind]
    torch.manual_seed(17)
    dec = make().to(dev).requires_grad_(False)
    for p in dec.parameters():
        p.normal_(0, 0.6)
    B, H, nrr, Sc, Sf = 2, 48, 16, 16, 16
    planes = torch.randn(B, 3, 32, H, H, device=dev)
    g = load_golden('renderer_seg')
    c2w = torch.from_numpy(g['cam2world'][:1]).to(dev).repeat(B, 1, 1)
    K = torch.from_numpy(g['intrinsics'][:1]).to(dev).repeat(B, 1, 1)
    o, d = native.ray_sampler(c2w, K, nrr)
    opts = dict(depth_resolution=Sc, depth_resolution_importance=Sf, ray_start=2.25, ray_end=3.3, box_warp=1.0,
                disparity_space_sampling=False, clamp_mode='softplus', white_back=False)
    monkeypatch.setattr(native, 'render_impl', impl)
    outs = []
    for fused in (True, False):
        r = ImportanceRenderer().to(dev)
        torch.manual_seed(99)                       # same torch.rand stream: jitter, then u
        if fused:
            with torch.no_grad():
                outs.append(r(planes, dec, o, d, opts))
        else:
            # gradients required -> the staged path; decoder kernels off -> the decoder MODULE's torch forward between the stages
            monkeypatch.setattr(native, 'decoder_mlp_supported', lambda *a, **k: False)
            outs.append(tuple(t.detach() for t in r(planes.clone().requires_grad_(True), dec, o, d, opts)))
    for a, b_, name in zip(outs[0], outs[1], ('features', 'depth', 'weights')):
        assert a.shape == b_.shape, name
        assert rel_err(a.cpu().numpy(), b_.cpu().numpy()) < 1e-3, (kind, name)
