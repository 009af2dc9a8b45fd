"""Time ImportanceRenderer.run_model at free points (the extract_mesh / G.sample_mixed query) on the GPU:
tensor-core query (p3d_run_model_tc) vs the CUDA-core kernel (p3d_run_model), dense 256^2 planes.

    python tools/time_query.py [--points 2097152] [--nets 2]
"""
import argparse
import json
import sys

import torch

sys.path.insert(0, '.')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--points', type=int, default=1 << 21)
    ap.add_argument('--nets', type=int, default=2)
    ap.add_argument('--iters', type=int, default=10)
    args = ap.parse_args()
    from pix2pix3d_b200 import native
    from pix2pix3d_b200.training.triplane import OSGDecoder
    from pix2pix3d_b200.training.triplane_cond import OSGDecoder_semantic_lateSeparate
    dev = torch.device('cuda')
    torch.manual_seed(0)
    if args.nets == 2:
        dec = OSGDecoder_semantic_lateSeparate(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32, 'sigmoid': False, 'semantic_channels': 6})
    else:
        dec = OSGDecoder(32, {'decoder_lr_mul': 1, 'decoder_output_dim': 32})
    packed = native.pack_decoder(dec.to(dev).requires_grad_(False))
    planes = torch.randn(1, 3, 256, 256, 32, device=dev)
    out = {}
    for pattern in ('grid', 'random'):
        if pattern == 'grid':
            n = round(args.points ** (1 / 3))
            ax = torch.linspace(-0.5, 0.5, n, device=dev)
            pts = torch.stack(torch.meshgrid(ax, ax, ax, indexing='ij'), -1).reshape(1, -1, 3)
        else:
            pts = torch.rand(1, args.points, 3, device=dev) - 0.5
        for impl in ('tc', 'simt'):
            for _ in range(3):
                native.run_model(planes, packed, pts, 1.0, impl=impl)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(args.iters):
                native.run_model(planes, packed, pts, 1.0, impl=impl)
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / args.iters
            out[f'{pattern}_{impl}'] = {'points': pts.shape[1], 'ms': round(ms, 4), 'Mpoints_per_s': round(pts.shape[1] / ms / 1e3, 1)}
        for _ in range(3):
            native.run_model(planes, packed, pts, 1.0, impl='tc', sigma_only=True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        ev[0].record()
        for _ in range(args.iters):
            native.run_model(planes, packed, pts, 1.0, impl='tc', sigma_only=True)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / args.iters
        out[f'{pattern}_tc_sigma_only'] = {'ms': round(ms, 4), 'Mpoints_per_s': round(pts.shape[1] / ms / 1e3, 1)}
    print(json.dumps({'nets': args.nets, **out}))


if __name__ == '__main__':
    main()
