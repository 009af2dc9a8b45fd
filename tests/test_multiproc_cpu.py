"""N > 1 host logic on CPU with the gloo backend (world_size 2): batch sharding of the render path reproduces the
unsharded result, and the gather of output images is the only collective involved."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden, rel_err
from make_golden import SYNTH_CASES, build_generator


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, q, limit=None):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import pix2pix3d_b200.training.triplane_cond as tc
        from pix2pix3d_b200 import sharding
        case = SYNTH_CASES[name]
        g = load_golden('synthesis_' + name)
        G = build_generator(tc, case)
        ws, c = torch.from_numpy(g['ws']), torch.from_numpy(g['c'])
        if limit is not None:
            ws, c = ws[:limit], c[:limit]
        B = ws.shape[0]
        lo, hi = sharding.shard_bounds(B, rank, world)
        nrr = case['nrr']
        # replay this shard's slice of the reference's renderer noise
        jit = torch.from_numpy(g['jitter'])[:B][lo:hi]
        u = torch.from_numpy(g['u']).reshape(-1, nrr * nrr, g['u'].shape[-1])[:B][lo:hi].reshape((hi - lo) * nrr * nrr, g['u'].shape[-1])
        it = iter([jit, u])
        o_like, o_rand = torch.rand_like, torch.rand
        torch.rand_like = lambda x, *a, **k: next(it)
        torch.rand = lambda *a, **k: next(it)
        try:
            with torch.no_grad():
                out = sharding.render_sharded(G, ws, c, gather=True, noise_mode='const', neural_rendering_resolution=nrr)
        finally:
            torch.rand_like, torch.rand = o_like, o_rand
        if rank == 0:
            q.put({k: v.numpy() for k, v in out.items()})
    finally:
        dist.destroy_process_group()


def test_sharded_render_matches_unsharded_reference():
    name = 'seg_tiny'          # batch of 2 -> one image per rank
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=240)
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    g = load_golden('synthesis_' + name)
    for k in ('image', 'semantic', 'image_raw', 'semantic_raw'):
        assert out[k].shape == g['out_' + k].shape
        assert rel_err(out[k], g['out_' + k]) < 1e-4, k
    # depth is clamped to the LOCAL batch's depth range (ray_marcher.py:50): identical unless the clamp is active
    assert rel_err(out['image_depth'], g['out_image_depth']) < 1e-3


def test_gather_with_an_empty_shard_does_not_deadlock():
    """Batch of 1 over 2 ranks: rank 1 renders nothing but still takes part in the gather."""
    name = 'seg_tiny'
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q, 1)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=240)
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    g = load_golden('synthesis_' + name)
    for k in ('image', 'semantic', 'image_raw', 'semantic_raw'):
        assert out[k].shape == g['out_' + k][:1].shape
        assert rel_err(out[k], g['out_' + k][:1]) < 1e-4, k


def test_shard_bounds_cover_batch_exactly():
    from pix2pix3d_b200 import sharding
    for n in (0, 1, 4, 7, 32):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
