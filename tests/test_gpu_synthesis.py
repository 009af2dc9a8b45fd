"""G.synthesis of the host-side mirror on CUDA (fused renderer + native ops) against the reference fixtures."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from make_golden import SEMGEN_CASES, SYNTH_CASES, build_generator, state_digest

pytestmark = pytest.mark.gpu


def _replay(g, dev):
    it = iter([torch.from_numpy(g['jitter']).to(dev), torch.from_numpy(g['u']).to(dev)])
    return (lambda x, *a, **k: next(it)), (lambda *a, **k: next(it))


@pytest.mark.parametrize('name', list(SYNTH_CASES) + list(SEMGEN_CASES))
@pytest.mark.parametrize('force_fp32', [True, False])
def test_synthesis_cuda_matches_reference(name, force_fp32):
    import pix2pix3d_b200.training.triplane_cond as tc
    from pix2pix3d_b200 import _lib
    case = SYNTH_CASES.get(name) or SEMGEN_CASES[name]
    g = load_golden('synthesis_' + name)
    dev = torch.device('cuda')
    torch.backends.cudnn.allow_tf32 = False        # any ATen convolution left on the path must be true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    G = build_generator(tc, case)
    assert state_digest(G) == bytes(g['state_digest']).decode()
    G = G.to(dev)
    ws, c = torch.from_numpy(g['ws']).to(dev), torch.from_numpy(g['c']).to(dev)
    rl, rr = _replay(g, dev)
    o_like, o_rand = torch.rand_like, torch.rand
    torch.rand_like, torch.rand = rl, rr
    before = _lib.launch_count
    try:
        with torch.no_grad():
            out = G.synthesis(ws, c, noise_mode='const', neural_rendering_resolution=case['nrr'], force_fp32=force_fp32)
    finally:
        torch.rand_like, torch.rand = o_like, o_rand
    assert _lib.launch_count > before, 'native kernels were not used'
    if name == 'seg_nrr64':
        assert _lib.launch_count - before > 40, 'whole-generator tensor-core path was expected here'
    # renderer outputs are fp32 in both modes; the SR stacks run fp16 unless force_fp32 (superresolution.py:304)
    for k in ('image_raw', 'image_depth', 'semantic_raw'):
        if k in out:
            assert rel_err(out[k].float().cpu().numpy(), g['out_' + k]) < 1e-3, k
    tol = 1e-3 if force_fp32 else 2e-2
    for k in ('image', 'semantic'):
        if k in out:
            assert out[k].dtype == torch.float32
            assert rel_err(out[k].cpu().numpy(), g['out_' + k]) < tol, (k, force_fp32)
    with torch.no_grad():
        smp = G.sample_mixed(torch.from_numpy(g['pts']).to(dev), None, ws, noise_mode='const')
    assert rel_err(smp['rgb'].cpu().numpy(), g['sample_rgb']) < 1e-3
    assert rel_err(smp['sigma'].cpu().numpy(), g['sample_sigma']) < 1e-3
    if 'semantic' in smp:
        assert rel_err(smp['semantic'].cpu().numpy(), g['sample_semantic']) < 1e-3


def test_mapping_cuda_matches_reference():
    import pix2pix3d_b200.training.triplane_cond as tc
    case = SYNTH_CASES['seg_tiny']
    g = load_golden('synthesis_seg_tiny')
    G = build_generator(tc, case).cuda()
    z, c, mask = (torch.from_numpy(g[k]).cuda() for k in ('z', 'c', 'mask'))
    with torch.no_grad():
        ws = G.mapping(z, c, {'mask': mask, 'pose': c})
    assert rel_err(ws.cpu().numpy(), g['ws']) < 1e-3


def _replayed(fn, g, dev):
    it = iter([torch.from_numpy(g['jitter']).to(dev), torch.from_numpy(g['u']).to(dev)])
    o_like, o_rand = torch.rand_like, torch.rand
    torch.rand_like, torch.rand = (lambda x, *a, **k: next(it)), (lambda *a, **k: next(it))
    try:
        return fn()
    finally:
        torch.rand_like, torch.rand = o_like, o_rand


def test_training_step_gradients_cuda_match_cpu():
    """BASELINE config 5 shape of work on a tiny model: G.synthesis WITH gradients (the op-by-op path: libp3d bias_act /
    upfirdn2d kernels under autograd, ATen convolutions and grid_sample), a non-saturating G loss through the
    DualDiscriminator, and the R1 double backward on D; gradients on CUDA must match the CPU evaluation of the same modules
    (whose forward is pinned to the reference fixtures by test_mirror_cpu.py)."""
    import pix2pix3d_b200.training.triplane_cond as tc
    from pix2pix3d_b200.training.dual_discriminator import DualDiscriminator
    case = SYNTH_CASES['seg_tiny']
    g = load_golden('synthesis_seg_tiny')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    def run(dev):
        G = build_generator(tc, case).to(dev).train().requires_grad_(True)
        torch.manual_seed(5)
        D = DualDiscriminator(c_dim=25, img_resolution=128, img_channels=3, channel_base=1024, channel_max=16,
                              num_fp16_res=0, conv_clamp=None).to(dev).train().requires_grad_(True)
        ws = torch.from_numpy(g['ws']).to(dev).requires_grad_(True)
        c = torch.from_numpy(g['c']).to(dev)
        out = _replayed(lambda: G.synthesis(ws, c, noise_mode='const', neural_rendering_resolution=case['nrr'], force_fp32=True),
                        g, torch.device(dev))
        img = {'image': out['image'], 'image_raw': out['image_raw']}
        logits = D(img, c)
        loss_g = torch.nn.functional.softplus(-logits).mean()
        names = ['backbone.synthesis.b4.conv1.weight', 'backbone.synthesis.b32.conv0.affine.bias', 'decoder.net.0.weight',
                 'superresolution.block1.conv1.weight', 'backbone.synthesis.b256.torgb.bias']
        params = dict(G.named_parameters())
        grads = torch.autograd.grad(loss_g, [ws] + [params[n] for n in names], retain_graph=False)
        # R1 on real images (double backward through D, incl. bias_act second order and upfirdn2d gradients)
        torch.manual_seed(6)
        real = {'image': torch.randn(2, 3, 128, 128).to(dev).requires_grad_(True),
                'image_raw': torch.randn(2, 3, 16, 16).to(dev).requires_grad_(True)}
        rl = D(real, c)
        r1 = torch.autograd.grad(rl.sum(), [real['image'], real['image_raw']], create_graph=True)
        pen = r1[0].square().sum([1, 2, 3]) + r1[1].square().sum([1, 2, 3])
        dparams = dict(D.named_parameters())
        dn = ['b128.conv0.weight', 'b8.conv1.bias', 'b4.out.weight']
        dgr = torch.autograd.grad(pen.mean(), [dparams[n] for n in dn])
        return ([loss_g.detach().cpu().numpy()] + [t.detach().cpu().numpy() for t in grads],
                [pen.detach().cpu().numpy()] + [t.detach().cpu().numpy() for t in dgr])

    from pix2pix3d_b200 import _lib
    cpu_g, cpu_d = run('cpu')
    before = _lib.launch_count
    gpu_g, gpu_d = run('cuda')
    assert _lib.launch_count > before, 'native op kernels were not used on the training path'
    for a, b in zip(gpu_g, cpu_g):
        assert rel_err(a, b) < 2e-3
    for a, b in zip(gpu_d, cpu_d):
        assert rel_err(a, b) < 2e-3
