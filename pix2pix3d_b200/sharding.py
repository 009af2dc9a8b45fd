"""Batch sharding of the render path across one process per GPU.

Every image (and every ray) of `G.synthesis` is independent, so inference shards the batch dimension of `(ws, c)` across
ranks with weights replicated and NO collective on the data path; a gather of the output images is optional (only if one
rank must hold them all). The reference does the same for training data (`InfiniteSampler(rank, num_replicas)`,
torch_utils/misc.py:113-144, training_loop.py:288) and has no multi-GPU inference at all.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous, balanced split of n items: rank r gets [lo, hi)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank=None, world=None):
    """Slice every tensor of a dict (or a single tensor) along dim 0 for this rank."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if isinstance(tensors, torch.Tensor):
        lo, hi = shard_bounds(tensors.shape[0], rank, world)
        return tensors[lo:hi]
    return {k: shard_batch(v, rank, world) for k, v in tensors.items()}


def gather_outputs(out, sizes=None):
    """all_gather a dict of per-rank output tensors along dim 0 (ragged shards allowed via `sizes`)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return out
    world = dist.get_world_size()
    res = {}
    for k, v in out.items():
        v = v.contiguous()
        if sizes is None:
            parts = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(parts, v)
        else:
            mx = max(sizes)
            pad = torch.zeros((mx,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
            pad[:v.shape[0]] = v
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad)
            parts = [p[:s] for p, s in zip(parts, sizes)]
        res[k] = torch.cat(parts, 0)
    return res


def render_sharded(G, ws, c, gather=False, **synthesis_kwargs):
    """Run `G.synthesis` on this rank's shard of (ws, c). With gather=True every rank returns the full batch."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(ws.shape[0], rank, world)
    out = G.synthesis(ws[lo:hi], c[lo:hi], **synthesis_kwargs) if hi > lo else {}
    if gather and world > 1:
        sizes = [shard_bounds(ws.shape[0], r, world)[1] - shard_bounds(ws.shape[0], r, world)[0] for r in range(world)]
        if min(sizes) == 0 and max(sizes) > 0:
            # a rank without images still has to enter every all_gather: it learns the output names / trailing shapes /
            # dtypes from the first non-empty rank and contributes zero-length tensors
            out = _empty_like_peer(out, sizes, ws.device)
        out = gather_outputs(out, sizes)
    return out


def _empty_like_peer(out, sizes, device):
    """Ranks with an empty shard receive (name, trailing shape, dtype) of every output from the first non-empty rank."""
    src = next(r for r, s in enumerate(sizes) if s > 0)
    meta = [[(k, tuple(v.shape[1:]), str(v.dtype).replace('torch.', '')) for k, v in out.items()]] if dist.get_rank() == src else [None]
    dist.broadcast_object_list(meta, src=src)
    if out:
        return out
    return {k: torch.zeros((0,) + shape, dtype=getattr(torch, dt), device=device) for k, shape, dt in meta[0]}
