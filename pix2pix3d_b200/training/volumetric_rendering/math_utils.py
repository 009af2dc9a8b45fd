"""Ray / box helpers (mirror of the reference's training/volumetric_rendering/math_utils.py)."""
import torch


def transform_vectors(matrix, vectors4):
    """Rows of `vectors4` [N,M] multiplied by `matrix` [M,M] from the left (reference :26)."""
    return torch.matmul(vectors4, matrix.T)


def normalize_vecs(vectors):
    return vectors / torch.norm(vectors, dim=-1, keepdim=True)


def torch_dot(x, y):
    return (x * y).sum(-1)


def get_ray_limits_box(rays_o, rays_d, box_side_length):
    """Slab test of rays against the axis-aligned cube of side `box_side_length` centred at the origin;
    returns (t_near, t_far) of shape [..., 1], (-1, -2) for misses (reference :46-98)."""
    if rays_o.device.type == 'cuda' and rays_o.dtype == torch.float32:
        from ... import native
        return native.ray_limits_box(rays_o, rays_d, box_side_length)
    lead = rays_o.shape[:-1]
    o = rays_o.detach().reshape(-1, 3)
    d = rays_d.detach().reshape(-1, 3)
    half = box_side_length / 2
    lo = torch.full([3], -half, dtype=o.dtype, device=o.device)
    hi = torch.full([3], half, dtype=o.dtype, device=o.device)
    inv = 1 / d
    neg = inv < 0
    # per axis: entry plane is `hi` for rays travelling in -axis direction, else `lo`
    t_in = (torch.where(neg, hi, lo) - o) * inv
    t_out = (torch.where(neg, lo, hi) - o) * inv
    valid = torch.ones(o.shape[0], dtype=torch.bool, device=o.device)
    tmin, tmax = t_in[:, 0], t_out[:, 0]
    for axis in (1, 2):
        valid = valid & ~((tmin > t_out[:, axis]) | (t_in[:, axis] > tmax))
        tmin = torch.max(tmin, t_in[:, axis])
        tmax = torch.min(tmax, t_out[:, axis])
    tmin = torch.where(valid, tmin, torch.full_like(tmin, -1))
    tmax = torch.where(valid, tmax, torch.full_like(tmax, -2))
    return tmin.reshape(*lead, 1), tmax.reshape(*lead, 1)


def linspace(start, stop, num):
    """[num, *start.shape] evenly spaced values from `start` to `stop` inclusive (reference :101-118)."""
    steps = torch.arange(num, dtype=torch.float32, device=start.device) / (num - 1)
    steps = steps.reshape([num] + [1] * start.ndim)
    return start[None] + steps * (stop - start)[None]
