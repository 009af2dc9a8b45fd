import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pix2pix3d_b200.torch_utils.ops import native_conv
dev = torch.device('cuda')
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
for (b, cin, cout, h, w, k) in [(4,64,64,32,32,3),(2,128,96,64,64,3),(2,128,64,64,64,3),(2,64,96,64,64,3),(2,128,128,64,64,3),(3,6,64,40,24,1),(3,64,64,40,24,1),(3,6,64,32,32,1),(2,128,3,64,64,1),(1,96,200,17,19,3),(1,128,128,17,19,3),(1,128,128,16,24,3)]:
    torch.manual_seed(0)
    x = torch.randn(b, cin, h, w, device=dev)
    wt = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    y = native_conv._conv(x, wt)
    yr = torch.nn.functional.conv2d(x.double(), wt.double(), padding=k // 2)
    gy = torch.randn_like(y)
    dx = native_conv._conv(gy, wt.flip([2, 3]).transpose(0, 1).contiguous())
    dxr = torch.nn.functional.conv_transpose2d(gy.double(), wt.double(), padding=k // 2)
    bad = (y.double() - yr).abs() > 1e-3
    print((b, cin, cout, h, w, k), 'fwd', f'{rel(y, yr):.2e}', 'dgrad', f'{rel(dx, dxr):.2e}', 'nan', bool(torch.isnan(y).any()), 'bad frac', float(bad.float().mean()),
          'bad channels', sorted(set(bad.nonzero()[:, 1].tolist()))[:8] if bad.any() else [])
