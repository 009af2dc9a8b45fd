"""Time the loss-side image ops of SURVEY 8f-4 against ATen at the config-5 shapes: the bilinear / anti-aliased resize
(p3d_resize_bilinear, forward and adjoint) and cross_entropy2d (p3d_cross_entropy2d_fwd/bwd vs the reference's
transpose + F.cross_entropy composition).

    python tools/time_loss_ops.py
"""
import json
import sys

import torch

sys.path.insert(0, '.')


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters * 1e3   # us


def main():
    from pix2pix3d_b200.torch_utils.ops.resize import _launch
    out = []
    for (b, c, src, dst) in ((4, 6, 128, 512), (4, 3, 512, 128), (4, 32, 64, 128), (32, 6, 128, 512)):
        x = torch.randn(b, c, src, src, device='cuda')
        g = torch.randn(b, c, dst, dst, device='cuda')
        nbytes = (x.numel() + g.numel()) * 4
        row = {'shape': [b, c, src, dst], 'MB': round(nbytes / 1e6, 2)}
        for name, fn in (
            ('p3d_fwd', lambda: _launch(x, (src, src), (dst, dst), True, False)),
            ('p3d_adj', lambda: _launch(g, (src, src), (dst, dst), True, True)),
            ('aten_fwd', lambda: torch.nn.functional.interpolate(x, size=(dst, dst), mode='bilinear', align_corners=False, antialias=True)),
            ('aten_adj', lambda: torch.ops.aten._upsample_bilinear2d_aa_backward(g, [dst, dst], [b, c, src, src], False, None, None)),
        ):
            us = timed(fn)
            row[name] = {'us': round(us, 2), 'GB/s': round(nbytes / us / 1e3, 1)}
        out.append(row)
    print(json.dumps({'resize': out}))
    from pix2pix3d_b200.training.loss_utils import cross_entropy2d
    ce = []
    for (b, c, r) in ((4, 6, 512), (4, 6, 128), (4, 19, 512)):
        x = torch.randn(b, c, r, r, device='cuda', requires_grad=True)
        t = torch.randint(0, c, (b, r, r), device='cuda')
        nbytes = x.numel() * 4 + t.numel() * 8

        def ref_fwd():
            flat = x.transpose(1, 2).transpose(2, 3).contiguous().view(-1, c)
            return torch.nn.functional.cross_entropy(flat, t.view(-1), reduction='mean')

        def both(fn):
            x.grad = None
            fn().backward()

        row = {'shape': [b, c, r], 'MB_fwd': round(nbytes / 1e6, 2)}
        with torch.no_grad():
            row['p3d_fwd_us'] = round(timed(lambda: cross_entropy2d(x, t)), 2)
            row['aten_fwd_us'] = round(timed(ref_fwd), 2)
        row['p3d_fwd_bwd_us'] = round(timed(lambda: both(lambda: cross_entropy2d(x, t))), 2)
        row['aten_fwd_bwd_us'] = round(timed(lambda: both(ref_fwd)), 2)
        row['p3d_fwd_GB/s'] = round(nbytes / row['p3d_fwd_us'] / 1e3, 1)
        ce.append(row)
    print(json.dumps({'cross_entropy2d': ce}))


if __name__ == '__main__':
    main()
