"""Time the training-path renderer stages at config-2 size (B=4, 128^2 rays, 96 samples): tri-plane lookup and ray marcher,
forward + backward, kernel-backed autograd Functions vs the ATen formulation.

    python tools/time_render_bwd.py
"""
import json
import sys

import torch

sys.path.insert(0, '.')


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(iters):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return round(ev[0].elapsed_time(ev[1]) / iters, 3)


def main():
    from pix2pix3d_b200.training.volumetric_rendering import ray_marcher as rm
    from pix2pix3d_b200.training.volumetric_rendering import renderer as rr
    dev = torch.device('cuda')
    torch.manual_seed(0)
    b, r, s, c = 4, 128 * 128, 96, 64
    out = {}
    colors = torch.rand(b, r, s, c, device=dev, requires_grad=True)
    dens = torch.randn(b, r, s, 1, device=dev, requires_grad=True)
    depths = torch.sort(torch.rand(b, r, s, 1, device=dev) + 2.25, dim=2)[0]
    g = torch.randn(b, r, c, device=dev)

    def march(fn):
        colors.grad = None; dens.grad = None
        rgb, depth, w = fn(colors, dens, depths, False)
        ((rgb * g).sum() + w.sum()).backward()

    out['ray_march_fwd_bwd_ms'] = {'p3d': timed(lambda: march(rm._RayMarch.apply)), 'aten': timed(lambda: march(rm._march_torch))}
    del colors, dens, depths, g
    torch.cuda.empty_cache()
    planes = torch.randn(b, 3, 32, 256, 256, device=dev, requires_grad=True)
    m = r * 48                                             # one pass of the renderer (coarse or fine)
    coords = torch.rand(b, m, 3, device=dev) - 0.5
    gf = torch.randn(b, 3, m, 32, device=dev)
    axes = rr.generate_planes().to(dev)

    def aten_sample(pf, co, bw):
        n, n_planes, cc, h, w = pf.shape
        grid = rr.project_onto_planes(axes, (2 / bw) * co).unsqueeze(1)
        o = torch.nn.functional.grid_sample(pf.reshape(n * n_planes, cc, h, w), grid, mode='bilinear', padding_mode='zeros', align_corners=False)
        return o.permute(0, 3, 2, 1).reshape(n, n_planes, co.shape[1], cc)

    def lookup(fn):
        planes.grad = None
        fn(planes, coords, 1.0).backward(gf)

    out['sample_from_planes_fwd_bwd_ms'] = {'p3d': timed(lambda: lookup(rr._SamplePlanes.apply)), 'aten': timed(lambda: lookup(aten_sample))}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
