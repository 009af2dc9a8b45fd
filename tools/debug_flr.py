"""GPU debug helper: fused filtered_lrelu vs the reference's stock plugin (signs, forward, backward) for one configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'baseline'))
import torch
import ref_harness as rh
rh.import_reference()
import torch_utils.ops.filtered_lrelu as r_fl
from pix2pix3d_b200.torch_utils import custom_ops as ours
from pix2pix3d_b200.torch_utils.ops import filtered_lrelu as my_fl
dev = torch.device('cuda')
assert r_fl._init()
stock = r_fl._plugin
mine = ours.get_plugin('filtered_lrelu_plugin')
import scipy.signal
f4 = torch.as_tensor(scipy.signal.firwin(numtaps=12, cutoff=0.25, width=0.3, fs=2.0), dtype=torch.float32, device=dev)   # separable: both plugins fuse
for dtype in (torch.float32, torch.float16):
    x = torch.randn(2, 6, 25, 20, device=dev, generator=torch.Generator(dev).manual_seed(1)).to(dtype)
    b = torch.randn(6, device=dev, generator=torch.Generator(dev).manual_seed(2)).to(dtype)
    args = (x, f4, f4, b, torch.empty(0, device=dev), 2, 2, 10, 9, 10, 9, 0, 0, 1.0, 0.2, 0.8, False, True)
    ys, sos, rcs = stock.filtered_lrelu(*args)
    ym, som, rcm = mine.filtered_lrelu(*args)
    print(dtype, 'rc', rcs, rcm, 'y shape', tuple(ys.shape), tuple(ym.shape), 'so shape', tuple(sos.shape), tuple(som.shape))
    print('  fwd max diff', (ys.float() - ym.float()).abs().max().item(), 'max', ys.float().abs().max().item())
    aw = ym.shape[3] * 2 - 1 + 11
    bs = torch.stack([(sos >> (2 * k)) & 3 for k in range(4)], -1).reshape(*sos.shape[:3], -1)[..., :aw]
    bm = torch.stack([(som >> (2 * k)) & 3 for k in range(4)], -1).reshape(*som.shape[:3], -1)[..., :aw]
    print('  sign records differing', (bs != bm).sum().item(), 'of', bs.numel())
    # backward with the STOCK signs on both plugins, and with own signs
    dy = torch.randn_like(ys)
    pp = (11 + 11 - 10, 20 * 2 - ym.shape[3] * 2 + 10 - 1, 11 + 11 - 10, 25 * 2 - ym.shape[2] * 2 + 10 - 1)
    bargs = lambda si: (dy, f4, f4, torch.zeros(6, device=dev, dtype=dtype), si, 2, 2, *[pp[0], pp[1], pp[2], pp[3]], 0 - 11 + 10, 0 - 11 + 10, 1.0 * 4 / 4, 0.2, float('inf'), True, False)
    gs, _, r1 = stock.filtered_lrelu(*bargs(sos))
    gm, _, r2 = mine.filtered_lrelu(*bargs(sos))
    gm2, _, r3 = mine.filtered_lrelu(*bargs(som))
    print('  bwd rc', r1, r2, r3, 'stock-signs diff', (gs.float() - gm.float()).abs().max().item(), 'own-signs diff', (gs.float() - gm2.float()).abs().max().item(), 'max', gs.float().abs().max().item())
    d = (gs.float() - gm.float()).abs()
    idx = (d > 1e-3).nonzero()
    print('  bad elements', idx.shape[0], idx[:8].tolist())
