"""The BASELINE.json workloads as constructor arguments + synthetic inputs.

Generator kwargs follow what the reference's train.py assembles (train.py:287-288, 314-316, 343-353, 374-383,
404-484): z_dim = w_dim = 512, c_dim = 25, channel_base 32768, channel_max 512, mapping num_layers 2,
fused_modconv_default 'inference_only', backbone fp32 (num_fp16_res 0), super-resolution fp16 (sr_num_fp16_res 4).
No checkpoint or dataset is available offline, so weights are seeded random and inputs synthetic.
"""
import math

import numpy as np
import torch

from . import camera_utils

RENDER_PRESETS = {
    # train.py:425-461
    'afhq': dict(depth_resolution=48, depth_resolution_importance=48, ray_start=2.25, ray_end=3.3, box_warp=1,
                 avg_camera_radius=2.7, avg_camera_pivot=[0, 0, -0.06]),
    'celeba': dict(depth_resolution=48, depth_resolution_importance=48, ray_start=2.25, ray_end=3.3, box_warp=1,
                   avg_camera_radius=2.7, avg_camera_pivot=[0, 0, 0.2]),
    'shapenet': dict(depth_resolution=64, depth_resolution_importance=64, ray_start=0.1, ray_end=2.6, box_warp=1.6,
                     white_back=True, avg_camera_radius=1.7, avg_camera_pivot=[0, 0, 0]),
}

WORKLOADS = {
    # name: (render preset, img_resolution, semantic_channels, neural rendering res, label kind, label res, batch)
    'seg2cat_smoke': dict(preset='afhq', img_resolution=128, semantic_channels=6, nrr=64, label='mask', label_res=64, batch=1,
                          depth_resolution=24, depth_resolution_importance=24),                    # configs[0]
    'seg2cat_512': dict(preset='afhq', img_resolution=512, semantic_channels=6, nrr=128, label='mask', label_res=512, batch=4),  # configs[1]
    'seg2face_512': dict(preset='celeba', img_resolution=512, semantic_channels=19, nrr=128, label='mask', label_res=512, batch=16),
    'edge2car_128': dict(preset='shapenet', img_resolution=128, semantic_channels=1, nrr=64, label='edge', label_res=128, batch=8),
}

_SR = {512: ('SuperresolutionHybrid8XDC', 'SuperresolutionHybrid8XDC_semantic'),
       128: ('SuperresolutionHybrid2X', 'SuperresolutionHybrid2X_semantic')}


def generator_kwargs(name):
    w = WORKLOADS[name]
    sr, sr_sem = _SR[w['img_resolution']]
    rk = dict(image_resolution=w['img_resolution'], disparity_space_sampling=False, clamp_mode='softplus',
              superresolution_module='training.superresolution.' + sr,
              superresolution_module_semantic='training.superresolution.' + sr_sem,
              c_gen_conditioning_zero=False, gpc_reg_prob=0.5, c_scale=1.0, superresolution_noise_mode='none',
              density_reg=0.25, density_reg_p_dist=0.004, reg_type='l1', decoder_lr_mul=1.0, sr_antialias=True)
    rk.update(RENDER_PRESETS[w['preset']])
    for k in ('depth_resolution', 'depth_resolution_importance'):
        if k in w:
            rk[k] = w[k]
    mapping = dict(class_name='training.triplane_cond.' + ('MaskMappingNetwork_disentangle' if w['label'] == 'mask'
                                                           else 'EdgeMappingNetwork_disentangle'),
                   num_layers=2, in_resolution=w['label_res'], in_channels=w['semantic_channels'] if w['label'] == 'mask' else 1)
    return dict(z_dim=512, c_dim=25, w_dim=512, img_resolution=w['img_resolution'], img_channels=3,
                semantic_channels=w['semantic_channels'], mapping_kwargs=mapping, rendering_kwargs=rk,
                channel_base=32768, channel_max=512, fused_modconv_default='inference_only', num_fp16_res=0,
                sr_num_fp16_res=4, conv_clamp=None,
                sr_kwargs=dict(channel_base=32768, channel_max=512, fused_modconv_default='inference_only'))


def build_generator(name, seed=0, device='cpu', with_mapping=True):
    """Seeded random-init TriPlaneSemanticEntangleGenerator for a workload; noise strengths are made non-zero so the
    noise path is live (they initialise to 0, networks_stylegan2.py:310)."""
    from .training import triplane_cond
    kw = generator_kwargs(name)
    if not with_mapping:   # the synthesis benchmark never calls mapping(): skip the 52 M-parameter label encoder
        kw['mapping_kwargs'] = dict(class_name='training.networks_stylegan2.MappingNetwork', num_layers=2)
    torch.manual_seed(seed)
    G = triplane_cond.TriPlaneSemanticEntangleGenerator(**kw).eval().requires_grad_(False)
    g = torch.Generator().manual_seed(seed + 1)
    for pname, p in G.named_parameters():
        if pname.endswith('noise_strength'):
            p.copy_(torch.randn([], generator=g) * 0.1)
    return G.to(device)


def camera_labels(batch, seed, preset='afhq'):
    """c [B,25]: look-at poses at yaw pi/2 +- 0.35, pitch pi/2 +- 0.25 (generate_video.py:58-61) | intrinsics."""
    r = RENDER_PRESETS[preset]
    rng = np.random.RandomState(seed)
    fov = 18.837 if preset != 'shapenet' else 45.0
    rows = []
    for _ in range(batch):
        c2w = camera_utils.LookAtPoseSampler.sample(math.pi / 2 + rng.uniform(-0.35, 0.35), math.pi / 2 + rng.uniform(-0.25, 0.25),
                                                    torch.tensor(r['avg_camera_pivot'], dtype=torch.float32),
                                                    radius=r['avg_camera_radius'])
        rows.append(torch.cat([c2w.reshape(1, 16), camera_utils.FOV_to_intrinsics(fov).reshape(1, 9)], 1))
    return torch.cat(rows, 0)


def synthetic_ws(batch, num_ws, seed):
    """W+ latents of the shape G.mapping returns; statistically like mapped latents (unit-variance rows)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, num_ws, 512, generator=g)


def label_map(name, batch, seed):
    w = WORKLOADS[name]
    g = torch.Generator().manual_seed(seed)
    res = w['label_res']
    if w['label'] == 'mask':
        blocks = torch.randint(0, w['semantic_channels'], (batch, 1, 32, 32), generator=g)
        return blocks.repeat_interleave(res // 32, 2).repeat_interleave(res // 32, 3).to(torch.uint8)
    return (torch.rand(batch, 1, res, res, generator=g) < 0.05).float() * 2 - 1


def render_algorithmic_bytes(batch, rays, samples):
    """touched bytes of the fused renderer = samples * (3 planes * 4 taps * 32 ch * 4 B) + ray I/O."""
    return batch * rays * samples * 1536 + batch * rays * (24 + 264)
