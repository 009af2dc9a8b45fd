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
        if (colors.device.type == 'cuda' and not depths.requires_grad and colors.dtype == torch.float32
                and colors.shape[2] <= 256 and colors.shape[3] <= 256):
            return _RayMarch.apply(colors, densities, depths, white_back)     # training: kernels in both directions
        return _march_torch(colors, densities, depths, white_back)

    def forward(self, colors, densities, depths, rendering_options):
        return self.run_forward(colors, densities, depths, rendering_options)


class _RayMarch(torch.autograd.Function):
    """p3d_ray_march forward, p3d_ray_march_bwd backward (first order; depths are not differentiated: in the reference they
    come from the stratified / importance samplers, which carry no gradient)."""

    @staticmethod
    def forward(ctx, colors, densities, depths, white_back):
        ctx.set_materialize_grads(False)
        ctx.white_back = white_back
        ctx.save_for_backward(colors, densities, depths)
        return native.ray_march(colors.detach(), densities.detach(), depths, white_back)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_rgb, g_depth, g_weights):
        colors, densities, depths = ctx.saved_tensors
        if g_rgb is None:
            g_rgb = colors.new_zeros(colors.shape[0], colors.shape[1], colors.shape[3])
        rng = None
        if g_depth is not None:
            lo, hi = torch.aminmax(depths)
            rng = torch.stack([lo, hi]).float()
        g_col, g_den = native.ray_march_bwd(colors, densities, depths, g_rgb, g_depth, g_weights, rng, ctx.white_back)
        return g_col.to(colors.dtype), g_den.to(densities.dtype), None, None


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
