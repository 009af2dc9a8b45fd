"""A/B timing of the FIR formulations (p3d_fir_act_nhwc_variant): python tools/time_fir.py [reps] [only_variant]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pix2pix3d_b200 import tcconv
from pix2pix3d_b200.torch_utils.ops import upfirdn2d
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
only = int(sys.argv[2]) if len(sys.argv) > 2 else None
f = upfirdn2d.setup_filter([1, 3, 3, 1]).cuda()
shapes = [('sr512 fp16', torch.float16, 1, 4, 512, 128), ('sr256 fp16', torch.float16, 1, 4, 256, 256), ('b256 fp32', torch.float32, 2, 4, 256, 128)]
for name, dt, planes, b, res, c in shapes:
    x = torch.randn(b, res + 1, res + 1, c, device='cuda').to(dt)
    noise = torch.randn(res, res, device='cuda')
    bias = torch.randn(c, device='cuda')
    ref = None
    for v in (1, 3):
        if only is not None and v != only:
            continue
        tcconv.FIR_VARIANT = v
        y = tcconv.fir_act_nhwc(x, f, noise, bias, planes, (res, res), act_gain=1.4142135, clamp=256.0)
        if ref is None:
            ref = y
        same = bool(torch.equal(ref, y))
        diff = float((ref.float() - y.float()).abs().max()), float((ref != y).float().mean())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            tcconv.fir_act_nhwc(x, f, noise, bias, planes, (res, res), act_gain=1.4142135, clamp=256.0)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        byts = x.numel() * x.element_size() + y.numel() * 2
        print(f'{name:12s} variant {v}: {us:8.1f} us  {byts / us / 1e3:7.1f} GB/s  identical={same} maxdiff={diff[0]:.3g} frac_diff={diff[1]:.3g}', flush=True)
