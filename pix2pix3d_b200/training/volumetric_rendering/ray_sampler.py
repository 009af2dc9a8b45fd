"""Camera -> per-pixel ray origins and directions.

Mirror of the reference's training/volumetric_rendering/ray_sampler.py:24-62 (OpenCV camera convention,
pixel centres, ray m = row * resolution + col). CUDA inputs run `p3d_ray_sampler` (one launch instead of ~15).
"""
import torch

from ... import native


class RaySampler(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.ray_origins_h, self.ray_directions, self.depths, self.image_coords, self.rendering_options = None, None, None, None, None

    def forward(self, cam2world_matrix, intrinsics, resolution):
        """cam2world_matrix [N,4,4], intrinsics [N,3,3] (normalised) -> origins [N,M,3], directions [N,M,3]."""
        if cam2world_matrix.device.type == 'cuda' and not (cam2world_matrix.requires_grad or intrinsics.requires_grad):
            return native.ray_sampler(cam2world_matrix, intrinsics, int(resolution))
        return _ray_sampler_torch(cam2world_matrix, intrinsics, resolution)


def _ray_sampler_torch(cam2world, K, res):
    n = cam2world.shape[0]
    dev = cam2world.device
    cam_loc = cam2world[:, :3, 3]
    fx, fy, cx, cy, sk = (K[:, 0, 0, None], K[:, 1, 1, None], K[:, 0, 2, None], K[:, 1, 2, None], K[:, 0, 1, None])
    ticks = torch.arange(res, dtype=torch.float32, device=dev) * (1. / res) + (0.5 / res)
    y_cam = ticks.repeat_interleave(res)[None].expand(n, -1)   # row index is the slow axis
    x_cam = ticks.repeat(res)[None].expand(n, -1)
    z_cam = torch.ones((n, res * res), device=dev)
    x_lift = (x_cam - cx + cy * sk / fy - sk * y_cam / fy) / fx * z_cam
    y_lift = (y_cam - cy) / fy * z_cam
    pts = torch.stack((x_lift, y_lift, z_cam, torch.ones_like(z_cam)), dim=-1)
    world = torch.bmm(cam2world, pts.permute(0, 2, 1)).permute(0, 2, 1)[:, :, :3]
    dirs = torch.nn.functional.normalize(world - cam_loc[:, None, :], dim=2)
    origins = cam_loc.unsqueeze(1).repeat(1, dirs.shape[1], 1)
    return origins, dirs
