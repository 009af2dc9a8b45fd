"""Midpoint-rule alpha compositing along rays.

Mirror of the reference's training/volumetric_rendering/ray_marcher.py:25-62 (`MipRayMarcher2`). CUDA inputs
that need no gradient run `p3d_ray_march`; other inputs use the torch formulation.
"""
import torch
import torch.nn.functional as F

from ... import native


class MipRayMarcher2(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def run_forward(self, colors, densities, depths, rendering_options):
        """colors [B,R,S,C], densities [B,R,S,1], depths [B,R,S,1] -> rgb [B,R,C], depth [B,R,1], weights [B,R,S-1,1]."""
        assert rendering_options['clamp_mode'] == 'softplus', "MipRayMarcher only supports `clamp_mode`=`softplus`!"
        white_back = bool(rendering_options.get('white_back', False))
        no_grad = not (torch.is_grad_enabled() and (colors.requires_grad or densities.requires_grad or depths.requires_grad))
        if colors.device.type == 'cuda' and no_grad:
            return native.ray_march(colors, densities, depths, white_back)
        return _march_torch(colors, densities, depths, white_back)

    def forward(self, colors, densities, depths, rendering_options):
        return self.run_forward(colors, densities, depths, rendering_options)


def _march_torch(colors, densities, depths, white_back):
    d0, d1 = depths[:, :, :-1], depths[:, :, 1:]
    delta = d1 - d0
    c_mid = (colors[:, :, :-1] + colors[:, :, 1:]) / 2
    s_mid = (densities[:, :, :-1] + densities[:, :, 1:]) / 2
    d_mid = (d0 + d1) / 2
    sigma = F.softplus(s_mid - 1)          # density activation with a -1 bias
    alpha = 1 - torch.exp(-(sigma * delta))
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :, :1]), 1 - alpha + 1e-10], -2), -2)[:, :, :-1]
    weights = alpha * trans
    rgb = torch.sum(weights * c_mid, -2)
    w_total = weights.sum(2)
    depth = torch.sum(weights * d_mid, -2) / w_total
    depth = torch.nan_to_num(depth, float('inf'))
    depth = torch.clamp(depth, torch.min(depths), torch.max(depths))
    if white_back:
        rgb = rgb + 1 - w_total
    rgb = rgb * 2 - 1
    return rgb, depth, weights
