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
        assert _lib.launch_count - before > 60, 'whole-generator tensor-core path was expected here'
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
