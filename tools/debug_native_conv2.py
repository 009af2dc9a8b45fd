import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import pix2pix3d_b200.training.triplane_cond as tc
from make_golden import SYNTH_CASES, build_generator
from conftest import load_golden
from pix2pix3d_b200.torch_utils.ops import native_conv, conv2d_gradfix
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
case = SYNTH_CASES['seg_nrr64']
g = load_golden('synthesis_seg_nrr64')
G = build_generator(tc, case).cuda().train().requires_grad_(True)
ws, c = torch.from_numpy(g['ws']).cuda(), torch.from_numpy(g['c']).cuda()
calls = []
orig = native_conv._conv
def logged(x, w):
    y = orig(x, w)
    yr = torch.nn.functional.conv2d(x.double(), w.double(), padding=w.shape[2] // 2)
    calls.append((tuple(x.shape), tuple(w.shape), float((y.double() - yr).abs().max() / yr.abs().max().clamp_min(1e-30)), x.is_contiguous(), w.is_contiguous()))
    return y
native_conv._conv = logged
grads = []
for on in (True, False):
    native_conv.enabled = on
    G.zero_grad(set_to_none=True)
    torch.manual_seed(5)
    out = G.synthesis(ws, c, noise_mode='const', neural_rendering_resolution=case['nrr'])
    (out['image'].square().mean() + out['semantic'].square().mean() + out['image_raw'].mean()).backward()
    grads.append({k: p.grad.detach().clone() for k, p in G.named_parameters() if p.grad is not None})
native_conv.enabled = True
for c_ in calls:
    print('conv', c_)
ga, gb = grads
errs = sorted(((float((ga[k] - gb[k]).abs().max() / gb[k].abs().max().clamp_min(1e-30)), k) for k in ga), reverse=True)
for e, k in errs[:25]:
    print(f'{e:.3e} {k} {tuple(ga[k].shape)}')
