"""numpy restatement of the synthesis side of training/networks_stylegan2.py, training/superresolution.py and
TriPlane*Generator.synthesis (training/triplane_cond.py:661-697, 1020-1061), evaluated from a state_dict
(name -> array) exactly as the reference's modules lay their parameters out. fp32 throughout (the reference's
CPU behaviour: blocks force fp32 off-CUDA, networks_stylegan2.py:423-425)."""
import numpy as np

from . import ops
from . import renderer as R

f32 = np.float32
_SQ2 = float(np.sqrt(2))


def fc_layer(x, sd, prefix, activation='linear', lr_multiplier=1.0):
    """FullyConnectedLayer.forward (networks_stylegan2.py:111-127)."""
    w = np.asarray(sd[prefix + '.weight'], f32)
    b = sd.get(prefix + '.bias')
    w = w * f32(lr_multiplier / np.sqrt(w.shape[1]))
    if b is not None:
        b = np.asarray(b, f32)
        if lr_multiplier != 1:
            b = b * f32(lr_multiplier)
    y = np.asarray(x, f32) @ w.T
    if activation == 'linear' and b is not None:
        return (y + b).astype(f32)
    return ops.bias_act(y, b, act=activation)


def synthesis_layer(x, w, sd, prefix, up=1, noise_mode='const', fused_modconv=True, gain=1, conv_clamp=None,
                    resample_filter=None, noise=None):
    """SynthesisLayer.forward (networks_stylegan2.py:313-332). noise: explicit [N,1,H,W] draw for 'random'."""
    styles = fc_layer(w, sd, prefix + '.affine')          # bias_init=1 lives in the state_dict
    nz = None
    if prefix + '.noise_const' in sd:
        strength = f32(sd[prefix + '.noise_strength'])
        if noise_mode == 'random':
            nz = np.asarray(noise, f32) * strength
        elif noise_mode == 'const':
            nz = np.asarray(sd[prefix + '.noise_const'], f32) * strength
    weight = np.asarray(sd[prefix + '.weight'], f32)
    x = ops.modulated_conv2d(x, weight, styles, noise=nz, up=up, padding=weight.shape[-1] // 2,
                             resample_filter=resample_filter, flip_weight=(up == 1), fused_modconv=fused_modconv)
    clamp = conv_clamp * gain if conv_clamp is not None else None
    return ops.bias_act(x, sd[prefix + '.bias'], act='lrelu', gain=_SQ2 * gain, clamp=clamp)


def torgb_layer(x, w, sd, prefix, conv_clamp=None, fused_modconv=True):
    """ToRGBLayer.forward (networks_stylegan2.py:354-359)."""
    weight = np.asarray(sd[prefix + '.weight'], f32)
    styles = fc_layer(w, sd, prefix + '.affine') * f32(1 / np.sqrt(weight.shape[1] * weight.shape[2] ** 2))
    x = ops.modulated_conv2d(x, weight, styles, demodulate=False, fused_modconv=fused_modconv)
    return ops.bias_act(x, sd[prefix + '.bias'], clamp=conv_clamp)


def synthesis_block(x, img, ws, sd, prefix, upsample=True, noise_mode='const', fused_modconv=True, conv_clamp=None,
                    noises=None):
    """SynthesisBlock / SynthesisBlockNoUp.forward, 'skip' architecture (networks_stylegan2.py:419-463,
    superresolution.py:244-289). ws [N, num_conv+num_torgb, w_dim]."""
    filt = np.asarray(sd[prefix + '.resample_filter'], f32)
    first = (prefix + '.const') in sd
    wi = 0
    nz = list(noises) if noises is not None else [None, None]
    if first:
        c = np.asarray(sd[prefix + '.const'], f32)
        x = np.broadcast_to(c[None], (ws.shape[0],) + c.shape).copy()
        x = synthesis_layer(x, ws[:, wi], sd, prefix + '.conv1', noise_mode=noise_mode, fused_modconv=fused_modconv,
                            conv_clamp=conv_clamp, noise=nz[0]); wi += 1
    else:
        x = synthesis_layer(x, ws[:, wi], sd, prefix + '.conv0', up=2 if upsample else 1, noise_mode=noise_mode,
                            fused_modconv=fused_modconv, conv_clamp=conv_clamp, resample_filter=filt, noise=nz[0]); wi += 1
        x = synthesis_layer(x, ws[:, wi], sd, prefix + '.conv1', noise_mode=noise_mode, fused_modconv=fused_modconv,
                            conv_clamp=conv_clamp, noise=nz[1]); wi += 1
    if upsample and img is not None:
        img = ops.upsample2d(img, filt)
    y = torgb_layer(x, ws[:, wi], sd, prefix + '.torgb', conv_clamp=conv_clamp, fused_modconv=fused_modconv)
    img = y if img is None else (img + y).astype(f32)
    return x, img


def synthesis_network(ws, sd, prefix, img_resolution, noise_mode='const', fused_modconv=True, conv_clamp=256):
    """SynthesisNetwork.forward (networks_stylegan2.py:505-520)."""
    ws = np.asarray(ws, f32)
    x = img = None
    w_idx = 0
    res = 4
    while res <= img_resolution:
        bp = f'{prefix}.b{res}'
        num_conv = 1 if (bp + '.const') in sd else 2
        x, img = synthesis_block(x, img, ws[:, w_idx:w_idx + num_conv + 1], sd, bp, noise_mode=noise_mode,
                                 fused_modconv=fused_modconv, conv_clamp=conv_clamp)
        w_idx += num_conv
        res *= 2
    return img


def resize_matrix(in_size, out_size, antialias=True):
    """Dense [out, in] matrix of one axis of F.interpolate(mode='bilinear', align_corners=False, antialias=...):
    ATen `_compute_indices_min_size_weights_aa` (triangle filter, support max(scale, 1), window clipped to the image and
    renormalised) or, without anti-aliasing, `area_pixel_compute_source_index` + `guard_index_and_lambda` (two clamped taps)."""
    scale = in_size / out_size
    m = np.zeros((out_size, in_size), np.float64)
    for o in range(out_size):
        if antialias:
            support = max(scale, 1.0)
            center = scale * (o + 0.5)
            lo = max(int(center - support + 0.5), 0)
            hi = min(int(center + support + 0.5), in_size)
            idx = np.arange(lo, hi)
            wgt = np.clip(1 - np.abs((idx - center + 0.5) / support), 0, None)
            m[o, lo:hi] = wgt / wgt.sum()
        else:
            src = max(scale * (o + 0.5) - 0.5, 0.0)
            i0 = min(int(src), in_size - 1)
            lam = min(max(src - i0, 0.0), 1.0)
            m[o, i0] += 1 - lam
            m[o, min(i0 + 1, in_size - 1)] += lam
    return m.astype(f32)


def bilinear_resize(x, size, antialias=True):
    """F.interpolate(x, size, mode='bilinear', align_corners=False, antialias=...) on [N,C,H,W]
    (superresolution.py:315-319, dual_discriminator.py:86-102)."""
    x = np.asarray(x, f32)
    wy = resize_matrix(x.shape[2], size[0], antialias)
    wx = resize_matrix(x.shape[3], size[1], antialias)
    return np.einsum('oh,nchw,pw->ncop', wy, x, wx, optimize=True).astype(f32)


def bilinear_resize_adjoint(g, in_size, antialias=True):
    """Gradient of `bilinear_resize` w.r.t. its input: g [N,C,out_h,out_w] -> [N,C,in_h,in_w]."""
    g = np.asarray(g, f32)
    wy = resize_matrix(in_size[0], g.shape[2], antialias)
    wx = resize_matrix(in_size[1], g.shape[3], antialias)
    return np.einsum('oh,ncop,pw->nchw', wy, g, wx, optimize=True).astype(f32)


def bilinear_antialias_resize(x, size):
    return bilinear_resize(x, size, antialias=True)


SR_KINDS = {  # class name -> (input_resolution, block0 upsamples?, resize only if smaller?)
    'SuperresolutionHybrid8X': (128, True, False), 'SuperresolutionHybrid8XDC': (128, True, False),
    'SuperresolutionHybrid8XDC_semantic': (128, True, False), 'SuperresolutionHybrid4X': (128, False, True),
    'SuperresolutionHybrid2X': (64, False, False), 'SuperresolutionHybrid2X_semantic': (64, False, False),
}


def superresolution(rgb, x, ws, sd, prefix, kind, noise_mode='none', fused_modconv=True, use_fp16_clamp=True, return_raw=False):
    """Superresolution*.forward (superresolution.py:48-57, 312-323). conv_clamp is 256 when the module was built
    with sr_num_fp16_res > 0 (it is kept even when the block runs in fp32)."""
    in_res, up0, only_smaller = SR_KINDS[kind]
    ws = np.asarray(ws, f32)[:, -1:, :].repeat(3, axis=1)
    need = (x.shape[-1] < in_res) if only_smaller else (x.shape[-1] != in_res)
    raw = rgb
    if need:
        x = bilinear_antialias_resize(x, (in_res, in_res))
        rgb = bilinear_antialias_resize(rgb, (in_res, in_res))
    clamp = 256 if use_fp16_clamp else None
    x, rgb = synthesis_block(x, rgb, ws, sd, prefix + '.block0', upsample=up0, noise_mode=noise_mode,
                             fused_modconv=fused_modconv, conv_clamp=clamp)
    if not up0 and not need:
        # SynthesisBlockNoUp adds its ToRGB in place into the image it was handed (superresolution.py:283 `img.add_(y)`);
        # without a resize that image is the `feature_image[:, :3]` view which synthesis also returns as the raw image
        # (triplane_cond.py:1055-1061), so the returned raw image carries the block-0 ToRGB term.
        raw = rgb
    x, rgb = synthesis_block(x, rgb, ws, sd, prefix + '.block1', upsample=True, noise_mode=noise_mode,
                             fused_modconv=fused_modconv, conv_clamp=clamp)
    return (rgb, raw) if return_raw else rgb


def decoder_from_state_dict(sd, prefix, kind, semantic_sigmoid=False, lr_mul=1.0):
    nets = [dict(w1=sd[f'{prefix}.net.0.weight'], b1=sd[f'{prefix}.net.0.bias'],
                 w2=sd[f'{prefix}.net.2.weight'], b2=sd[f'{prefix}.net.2.bias'])]
    if kind == 'OSGDecoder_semantic_lateSeparate':
        nets.append(dict(w1=sd[f'{prefix}.net_semantic.0.weight'], b1=sd[f'{prefix}.net_semantic.0.bias'],
                         w2=sd[f'{prefix}.net_semantic.2.weight'], b2=sd[f'{prefix}.net_semantic.2.bias']))
    return dict(kind=kind, nets=nets, sigmoid=semantic_sigmoid, lr_mul=lr_mul)


def generator_synthesis(ws, c, sd, cfg, jitter, u, noise_mode='const'):
    """TriPlaneSemanticEntangleGenerator.synthesis / TriPlaneGenerator.synthesis with explicit renderer noise.
    cfg: dict(nrr, rendering_kwargs, semantic_channels (0 = no semantic branch), sr_kind, sr_kind_semantic,
    sr_fp16 (bool: module built with sr_num_fp16_res>0), fused_modconv)."""
    ws = np.asarray(ws, f32)
    c = np.asarray(c, f32)
    rk = cfg['rendering_kwargs']
    nrr = cfg['nrr']
    n = ws.shape[0]
    origins, dirs = R.ray_sampler(c[:, :16].reshape(-1, 4, 4), c[:, 16:25].reshape(-1, 3, 3), nrr)
    fused = cfg.get('fused_modconv', True)
    planes = synthesis_network(ws, sd, 'backbone.synthesis', 256 if 'plane_res' not in cfg else cfg['plane_res'],
                               noise_mode=noise_mode, fused_modconv=fused, conv_clamp=cfg.get('conv_clamp', None))
    planes = planes.reshape(n, 3, 32, planes.shape[-2], planes.shape[-1])
    cs = cfg.get('semantic_channels', 0)
    if cs > 0:
        dec = decoder_from_state_dict(sd, 'decoder', 'OSGDecoder_semantic_lateSeparate', semantic_sigmoid=(cs == 1),
                                      lr_mul=rk.get('decoder_lr_mul', 1))
    else:
        dec = decoder_from_state_dict(sd, 'decoder', 'OSGDecoder', lr_mul=rk.get('decoder_lr_mul', 1))
    m = nrr * nrr
    depths_coarse = R.sample_stratified(n, m, rk['ray_start'], rk['ray_end'], rk['depth_resolution'], jitter,
                                        rk.get('disparity_space_sampling', False))
    feats, depth, _ = R.importance_renderer(planes, dec, origins, dirs, depths_coarse, u, rk)
    fimg = np.ascontiguousarray(feats.transpose(0, 2, 1).reshape(n, feats.shape[-1], nrr, nrr))
    dimg = depth.transpose(0, 2, 1).reshape(n, 1, nrr, nrr)
    sr_noise = rk['superresolution_noise_mode']
    out = {'image_depth': dimg, 'planes': planes}
    if cs > 0:
        half = fimg.shape[1] // 2
        rgb_f, sem_f = fimg[:, :half], fimg[:, half:]
        out['image'], out['image_raw'] = superresolution(rgb_f[:, :3], rgb_f, ws, sd, 'superresolution', cfg['sr_kind'],
                                                         noise_mode=sr_noise, fused_modconv=fused,
                                                         use_fp16_clamp=cfg.get('sr_fp16', True), return_raw=True)
        out['semantic'], out['semantic_raw'] = superresolution(sem_f[:, :cs], sem_f, ws, sd, 'superresolution_semantic',
                                                               cfg['sr_kind_semantic'], noise_mode=sr_noise, fused_modconv=fused,
                                                               use_fp16_clamp=cfg.get('sr_fp16', True), return_raw=True)
    else:
        out['image'], out['image_raw'] = superresolution(fimg[:, :3], fimg, ws, sd, 'superresolution', cfg['sr_kind'],
                                                         noise_mode=sr_noise, fused_modconv=fused,
                                                         use_fp16_clamp=cfg.get('sr_fp16', True), return_raw=True)
    return out


def generator_synthesis_semantic(ws, c, sd, cfg, jitter, u, noise_mode='const'):
    """TriPlaneSemanticGenerator.synthesis (triplane_cond.py:772-822): ws[..., :w_dim] drives the texture backbone and the
    colour super-resolution, ws[..., w_dim:] the semantic backbone and the semantic super-resolution."""
    ws = np.asarray(ws, f32)
    c = np.asarray(c, f32)
    rk = cfg['rendering_kwargs']
    nrr, wd, cs = cfg['nrr'], cfg['w_dim'], cfg['semantic_channels']
    n = ws.shape[0]
    ws_t, ws_s = np.ascontiguousarray(ws[..., :wd]), np.ascontiguousarray(ws[..., wd:])
    origins, dirs = R.ray_sampler(c[:, :16].reshape(-1, 4, 4), c[:, 16:25].reshape(-1, 3, 3), nrr)
    fused = cfg.get('fused_modconv', True)
    res = cfg.get('plane_res', 256)
    pt = synthesis_network(ws_t, sd, 'backbone.synthesis', res, noise_mode=noise_mode, fused_modconv=fused,
                           conv_clamp=cfg.get('conv_clamp', None))
    ps = synthesis_network(ws_s, sd, 'backbone_semantic.synthesis', res, noise_mode=noise_mode, fused_modconv=fused,
                           conv_clamp=cfg.get('conv_clamp', None))
    pt = pt.reshape(n, 3, 32, pt.shape[-2], pt.shape[-1])
    ps = ps.reshape(n, 3, 32, ps.shape[-2], ps.shape[-1])
    lr = rk.get('decoder_lr_mul', 1)
    dec_t = decoder_from_state_dict(sd, 'decoder', 'OSGDecoder', lr_mul=lr)
    dec_s = decoder_from_state_dict(sd, 'decoder_semantic', 'OSGDecoder_semantic', semantic_sigmoid=(cs == 1), lr_mul=lr)
    m = nrr * nrr
    depths_coarse = R.sample_stratified(n, m, rk['ray_start'], rk['ray_end'], rk['depth_resolution'], jitter,
                                        rk.get('disparity_space_sampling', False))
    feats, depth, _ = R.importance_semantic_renderer(pt, ps, dec_t, dec_s, origins, dirs, depths_coarse, u, rk)
    fimg = np.ascontiguousarray(feats.transpose(0, 2, 1).reshape(n, feats.shape[-1], nrr, nrr))
    dimg = depth.transpose(0, 2, 1).reshape(n, 1, nrr, nrr)
    sr_noise = rk['superresolution_noise_mode']
    half = fimg.shape[1] // 2
    rgb_f, sem_f = fimg[:, :half], fimg[:, half:]
    out = {'image_depth': dimg, 'planes_texture': pt, 'planes_semantic': ps}
    out['image'], out['image_raw'] = superresolution(rgb_f[:, :3], rgb_f, ws_t, sd, 'superresolution', cfg['sr_kind'],
                                                     noise_mode=sr_noise, fused_modconv=fused,
                                                     use_fp16_clamp=cfg.get('sr_fp16', True), return_raw=True)
    out['semantic'], out['semantic_raw'] = superresolution(sem_f[:, :cs], sem_f, ws_s, sd, 'superresolution_semantic',
                                                           cfg['sr_kind_semantic'], noise_mode=sr_noise, fused_modconv=fused,
                                                           use_fp16_clamp=cfg.get('sr_fp16', True), return_raw=True)
    return out
