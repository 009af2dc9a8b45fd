"""Host-side mirror (pix2pix3d_b200.training / torch_utils) on CPU tensors against the reference fixtures:
checks the module wiring, parameter naming/initialisation order and the torch formulation that CPU inputs take."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from make_golden import SEMGEN_CASES, SYNTH_CASES, build_generator, capture_rand, state_digest


@pytest.mark.parametrize('name', list(SYNTH_CASES) + list(SEMGEN_CASES))
def test_synthesis_cpu_matches_reference(name):
    import pix2pix3d_b200.training.triplane_cond as tc
    case = SYNTH_CASES.get(name) or SEMGEN_CASES[name]
    g = load_golden('synthesis_' + name)
    G = build_generator(tc, case)
    assert state_digest(G) == bytes(g['state_digest']).decode()
    z, c, mask = (torch.from_numpy(g[k]) for k in ('z', 'c', 'mask'))
    draws = []
    with torch.no_grad():
        ws = G.mapping(z, c, {'mask': mask, 'pose': c})
        assert rel_err(ws.numpy(), g['ws']) < 1e-5
        # same seed-independent check of RNG consumption order: first rand_like (jitter), then rand (u)
        with capture_rand(draws):
            torch.manual_seed(0)
            G.synthesis(ws, c, noise_mode='const', neural_rendering_resolution=case['nrr'])
        assert [d[0] for d in draws] == ['rand_like', 'rand']
        assert tuple(draws[0][1].shape) == tuple(g['jitter'].shape) and tuple(draws[1][1].shape) == tuple(g['u'].shape)
        # replay the reference's noise
        it = iter([torch.from_numpy(g['jitter']), torch.from_numpy(g['u'])])
        o_like, o_rand = torch.rand_like, torch.rand
        torch.rand_like = lambda x, *a, **k: next(it)
        torch.rand = lambda *a, **k: next(it)
        try:
            out = G.synthesis(ws, c, noise_mode='const', neural_rendering_resolution=case['nrr'])
        finally:
            torch.rand_like, torch.rand = o_like, o_rand
        smp = G.sample_mixed(torch.from_numpy(g['pts']), None, ws, noise_mode='const')
    for k, v in out.items():
        assert rel_err(v.numpy(), g['out_' + k]) < 1e-4, k
    assert rel_err(smp['rgb'].numpy(), g['sample_rgb']) < 1e-5
    assert rel_err(smp['sigma'].numpy(), g['sample_sigma']) < 1e-5
    if 'semantic' in smp:
        assert rel_err(smp['semantic'].numpy(), g['sample_semantic']) < 1e-5


def test_ops_ref_paths_match_reference():
    from pix2pix3d_b200.torch_utils.ops import bias_act, conv2d_resample, upfirdn2d
    from pix2pix3d_b200.training.networks_stylegan2 import modulated_conv2d
    g = load_golden('ops')
    t = lambda k: torch.from_numpy(g[k])
    x, b = t('ba_x'), t('ba_b')
    for act in bias_act.activation_funcs:
        y = bias_act.bias_act(x, b, act=act)
        assert rel_err(y.numpy(), g[f'ba_{act}_d_y']) < 1e-6
        y = bias_act.bias_act(x, b, act=act, gain=1.7, clamp=0.9, alpha=0.3, impl='ref')
        assert rel_err(y.numpy(), g[f'ba_{act}_c_y']) < 1e-6
    f4 = upfirdn2d.setup_filter([1, 3, 3, 1])
    assert torch.equal(f4, t('up_f4'))
    assert torch.allclose(upfirdn2d.setup_filter([1, 2, 3, 4, 4, 3, 2, 1]), t('up_f8'))
    xs = t('up_x')
    assert rel_err(upfirdn2d.upfirdn2d(xs, f4, padding=[1, 1, 1, 1], gain=4).numpy(), g['up_post_tconv_y']) < 1e-6
    assert rel_err(upfirdn2d.upsample2d(xs, f4).numpy(), g['up_skip_up_y']) < 1e-6
    assert rel_err(upfirdn2d.downsample2d(xs, f4).numpy(), g['up_down2_y']) < 1e-6
    assert rel_err(upfirdn2d.upfirdn2d(xs, t('up_f35'), up=[3, 2], down=[2, 1], padding=[2, 0, -1, 3], gain=0.7,
                                       flip_filter=True).numpy(), g['up_odd_y']) < 1e-6
    xc, w3, w1, st, nz = t('mc_x'), t('mc_w3'), t('mc_w1'), t('mc_styles'), t('mc_noise16')
    assert rel_err(conv2d_resample.conv2d_resample(xc, w3, f=f4, up=2, padding=1, flip_weight=False).numpy(), g['cr_up2']) < 1e-6
    assert rel_err(conv2d_resample.conv2d_resample(xc, w3, f=f4, down=2, padding=1).numpy(), g['cr_down2']) < 1e-6
    assert rel_err(conv2d_resample.conv2d_resample(xc, w1, f=f4, down=2).numpy(), g['cr_1x1_down2']) < 1e-6
    assert rel_err(conv2d_resample.conv2d_resample(xc, w1, f=f4, up=2).numpy(), g['cr_1x1_up2']) < 1e-6
    for fused in (True, False):
        tag = 'f' if fused else 'n'
        y = modulated_conv2d(xc.clone(), w3, st, noise=nz, up=2, padding=1, resample_filter=f4, flip_weight=False, fused_modconv=fused)
        assert rel_err(y.numpy(), g[f'mc_up2_{tag}']) < 1e-5
        y = modulated_conv2d(xc.clone(), w1, st, demodulate=False, fused_modconv=fused)
        assert rel_err(y.numpy(), g[f'mc_torgb_{tag}']) < 1e-5


def test_alias_install_resolves_reference_import_paths():
    import pix2pix3d_b200
    pix2pix3d_b200.install()
    try:
        import training.volumetric_rendering.renderer as r1
        import pix2pix3d_b200.training.volumetric_rendering.renderer as r2
        from torch_utils.ops import bias_act, upfirdn2d, conv2d_resample, fma, conv2d_gradfix, grid_sample_gradfix  # noqa: F401
        import dnnlib
        assert r1 is r2
        assert dnnlib.util.construct_class_by_name(class_name='training.volumetric_rendering.ray_sampler.RaySampler') is not None
        for name in ('ImportanceRenderer', 'sample_from_planes', 'generate_planes', 'project_onto_planes'):
            assert hasattr(r1, name)
    finally:
        pix2pix3d_b200.uninstall()


def test_bias_act_second_order_on_cpu_autograd():
    """ref path stays differentiable to second order (R1 penalty path)."""
    from pix2pix3d_b200.torch_utils.ops import bias_act
    x = torch.randn(2, 3, 4, 4, dtype=torch.float64, requires_grad=True)
    b = torch.randn(3, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradgradcheck(lambda a, c: bias_act.bias_act(a, c, act='swish'), (x, b))


def test_ray_limits_box_matches_reference_semantics():
    from pix2pix3d_b200.training.volumetric_rendering import math_utils
    o = torch.tensor([[0., 0., 2.], [0., 0., 2.], [3., 3., 3.]])
    d = torch.tensor([[0., 0., -1.], [0., 1., 0.], [0.1, 0.2, 1.]])
    d = d / d.norm(dim=-1, keepdim=True)
    tmin, tmax = math_utils.get_ray_limits_box(o, d + 1e-9, box_side_length=1.0)
    assert abs(tmin[0].item() - 1.5) < 1e-5 and abs(tmax[0].item() - 2.5) < 1e-5
    assert tmin[1].item() == -1 and tmax[1].item() == -2
    assert tmin[2].item() == -1 and tmax[2].item() == -2


def test_filtered_lrelu_and_dual_discriminator_match_reference():
    from pix2pix3d_b200.torch_utils.ops import filtered_lrelu
    from pix2pix3d_b200.training.dual_discriminator import DualDiscriminator
    g = load_golden('ops')
    t = lambda k: torch.from_numpy(g[k])
    y = filtered_lrelu.filtered_lrelu(t('fl_x'), t('up_f4'), t('fl_fd'), t('fl_b'), up=2, down=2, padding=[3, 2, 3, 2], clamp=0.8)
    assert rel_err(y.numpy(), g['fl_up2_down2']) < 1e-6
    torch.manual_seed(31)
    D = DualDiscriminator(c_dim=25, img_resolution=64, img_channels=3, channel_base=1024, channel_max=32, mapping_kwargs={},
                          epilogue_kwargs={'mbstd_group_size': 2}).eval().requires_grad_(False)
    out = D({'image': t('dd_image'), 'image_raw': t('dd_image_raw')}, t('dd_c').clone())
    assert rel_err(out.numpy(), g['dd_logits']) < 1e-5


def test_copy_params_and_buffers_semantic_fallback(capsys):
    """reference torch_utils/misc.py:157-176: `*_semantic.*` tensors missing in the source are initialised from the
    same-named tensors without the `_semantic` suffix (resuming from an EG3D checkpoint)."""
    import torch
    from pix2pix3d_b200.torch_utils import misc

    class Src(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.superresolution = torch.nn.Linear(3, 2)

    class Dst(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.superresolution = torch.nn.Linear(3, 2)
            self.superresolution_semantic = torch.nn.Linear(3, 2)
            self.other = torch.nn.Linear(2, 2)

    src, dst = Src(), Dst()
    before = dst.other.weight.clone()
    with torch.no_grad():
        misc.copy_params_and_buffers(src, dst, require_all=False)
    assert torch.equal(dst.superresolution_semantic.weight, src.superresolution.weight)
    assert torch.equal(dst.superresolution.bias, src.superresolution.bias)
    assert torch.equal(dst.other.weight, before)
    assert 'other.weight not found in source module' in capsys.readouterr().out
