"""Look-at camera poses and pinhole intrinsics for the 25-float camera label `c = [cam2world(16) | K(9)]`.

Same call surface as the reference's camera_utils.py (`LookAtPoseSampler.sample` :69, `create_cam2world_matrix`
:118, `FOV_to_intrinsics` :140); used here only to build synthetic poses for tests and benchmarks.
"""
import math

import torch


def _unit(v):
    return v / torch.norm(v, dim=-1, keepdim=True)


def create_cam2world_matrix(forward_vector, origin):
    """[N,4,4] OpenCV-style camera-to-world: columns are (right, up, forward, position); y is up, no roll."""
    fwd = _unit(forward_vector)
    world_up = torch.zeros_like(fwd)
    world_up[:, 1] = 1
    right = -_unit(torch.linalg.cross(world_up, fwd, dim=-1))
    up = _unit(torch.linalg.cross(fwd, right, dim=-1))
    m = torch.zeros(fwd.shape[0], 4, 4, device=origin.device, dtype=fwd.dtype)
    m[:, :3, 0], m[:, :3, 1], m[:, :3, 2], m[:, :3, 3] = right, up, fwd, origin
    m[:, 3, 3] = 1
    return m


def _sphere_point(yaw, pitch, radius):
    """Position on the sphere for yaw/pitch in radians (pi/2, pi/2 = on the +z axis); pitch uses the
    reference's arccos(1 - 2 v/pi) area-uniform parametrisation."""
    pitch = torch.clamp(pitch, 1e-5, math.pi - 1e-5)
    phi = torch.arccos(1 - 2 * (pitch / math.pi))
    x = radius * torch.sin(phi) * torch.cos(math.pi - yaw)
    z = radius * torch.sin(phi) * torch.sin(math.pi - yaw)
    y = radius * torch.cos(phi)
    return torch.cat([x, y, z], dim=1)


class LookAtPoseSampler:
    @staticmethod
    def sample(horizontal_mean, vertical_mean, lookat_position, horizontal_stddev=0, vertical_stddev=0, radius=1,
               batch_size=1, device='cpu'):
        yaw = torch.randn((batch_size, 1), device=device) * horizontal_stddev + horizontal_mean
        pitch = torch.randn((batch_size, 1), device=device) * vertical_stddev + vertical_mean
        origins = _sphere_point(yaw, pitch, radius)
        return create_cam2world_matrix(lookat_position - origins, origins)


class UniformCameraPoseSampler:
    @staticmethod
    def sample(horizontal_mean, vertical_mean, horizontal_stddev=0, vertical_stddev=0, radius=1, batch_size=1, device='cpu'):
        yaw = (torch.rand((batch_size, 1), device=device) * 2 - 1) * horizontal_stddev + horizontal_mean
        pitch = (torch.rand((batch_size, 1), device=device) * 2 - 1) * vertical_stddev + vertical_mean
        origins = _sphere_point(yaw, pitch, radius)
        return create_cam2world_matrix(-origins, origins)


def FOV_to_intrinsics(fov_degrees, device='cpu'):
    """Normalised intrinsics (principal point at the image centre); the reference's constants are kept."""
    f = float(1 / (math.tan(fov_degrees * 3.14159 / 360) * 1.414))
    return torch.tensor([[f, 0, 0.5], [0, f, 0.5], [0, 0, 1]], device=device)
