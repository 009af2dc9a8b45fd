"""numpy restatement of training/volumetric_rendering/{ray_sampler,renderer,ray_marcher}.py and the OSG decoders.

All arithmetic is float32 with one rounding per elementary operation (numpy never contracts a*b+c into an FMA),
in the operand order of the reference's Python expressions. Reductions whose order ATen leaves unspecified
(sum of the pdf weights, cumsum, cumprod, channel sums) are evaluated strictly left to right; the CUDA kernels
use the same order for the index-bearing ones (importance sampling), so indices can be compared bit-exactly.
"""
import numpy as np

f32 = np.float32


# ---------------------------------------------------------------------------------------------
# ray_sampler.py:24-62
# ---------------------------------------------------------------------------------------------
def ray_sampler(cam2world, intrinsics, resolution):
    """cam2world [N,4,4], intrinsics [N,3,3] -> origins [N,M,3], dirs [N,M,3]; ray m = row*res + col."""
    c2w = np.asarray(cam2world, f32).reshape(-1, 4, 4)
    K = np.asarray(intrinsics, f32).reshape(-1, 3, 3)
    n, res = c2w.shape[0], int(resolution)
    fx, fy, cx, cy, sk = (K[:, 0, 0, None], K[:, 1, 1, None], K[:, 0, 2, None], K[:, 1, 2, None], K[:, 0, 1, None])
    ticks = np.arange(res, dtype=f32) * f32(1.0 / res) + f32(0.5 / res)          # :43
    x_cam = np.tile(ticks, res)[None].repeat(n, 0)                               # column index is fastest (:44)
    y_cam = np.repeat(ticks, res)[None].repeat(n, 0)
    x_lift = (x_cam - cx + cy * sk / fy - sk * y_cam / fy) / fx                  # :51 (z_cam == 1)
    y_lift = (y_cam - cy) / fy                                                   # :52
    world = np.empty((n, res * res, 3), f32)
    for i in range(3):                                                           # bmm with [x, y, 1, 1] (:56)
        acc = c2w[:, i, 0, None] * x_lift
        acc = acc + c2w[:, i, 1, None] * y_lift
        acc = acc + c2w[:, i, 2, None]
        acc = acc + c2w[:, i, 3, None]
        world[:, :, i] = acc
    cam_loc = c2w[:, :3, 3]
    d = world - cam_loc[:, None, :]                                              # :58
    nrm = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2], dtype=f32)
    d = d / np.maximum(nrm, f32(1e-12))[..., None]                               # F.normalize (:59)
    origins = np.broadcast_to(cam_loc[:, None, :], d.shape).copy()               # :61
    return origins.astype(f32), d.astype(f32)


# ---------------------------------------------------------------------------------------------
# renderer.py:169-192
# ---------------------------------------------------------------------------------------------
def torch_linspace(start, end, steps):
    """torch.linspace for float32: symmetric evaluation around the midpoint (ATen RangeFactories)."""
    start, end = f32(start), f32(end)
    step = f32((end - start) / f32(steps - 1))
    idx = np.arange(steps)
    lo = (start + step * idx.astype(f32)).astype(f32)
    hi = (end - step * (steps - 1 - idx).astype(f32)).astype(f32)
    return np.where(idx < steps // 2, lo, hi).astype(f32)


def sample_stratified(n, m, ray_start, ray_end, depth_resolution, jitter, disparity_space_sampling=False):
    """jitter: the `torch.rand_like` draw [N,M,S,1] (renderer.py:181/190). Returns depths [N,M,S,1]."""
    jitter = np.asarray(jitter, f32).reshape(n, m, depth_resolution, 1)
    if disparity_space_sampling:
        t = torch_linspace(0, 1, depth_resolution).reshape(1, 1, -1, 1) + jitter * f32(1 / (depth_resolution - 1))
        return (f32(1.) / (f32(1. / ray_start) * (f32(1.) - t) + f32(1. / ray_end) * t)).astype(f32)
    base = torch_linspace(ray_start, ray_end, depth_resolution).reshape(1, 1, -1, 1)
    delta = f32((ray_end - ray_start) / (depth_resolution - 1))
    return (base + jitter * delta).astype(f32)


# ---------------------------------------------------------------------------------------------
# renderer.py:23-65 (generate_planes / project_onto_planes / sample_from_planes)
# ---------------------------------------------------------------------------------------------
PLANE_COORDS = ((0, 1), (0, 2), (2, 0))   # plane k samples (x,y), (x,z), (z,x): inverse of the axes at :30-37


def _bilinear_zeros(plane, gx, gy):
    """F.grid_sample(bilinear, zeros, align_corners=False) of plane [C,H,W] at normalised (gx, gy) [M] -> [M,C]."""
    c, h, w = plane.shape
    ix = ((gx + f32(1)) * f32(w) - f32(1)) * f32(0.5)
    iy = ((gy + f32(1)) * f32(h) - f32(1)) * f32(0.5)
    x0f, y0f = np.floor(ix), np.floor(iy)
    ax, bx = (x0f + f32(1)) - ix, ix - x0f
    ay, by = (y0f + f32(1)) - iy, iy - y0f
    inside = (ix > -1) & (ix < w) & (iy > -1) & (iy < h)
    x0 = np.where(inside, x0f, 0).astype(np.int64)
    y0 = np.where(inside, y0f, 0).astype(np.int64)
    out = np.zeros((gx.shape[0], c), f32)
    for dy, dx, wgt in ((0, 0, ax * ay), (0, 1, bx * ay), (1, 0, ax * by), (1, 1, bx * by)):   # nw, ne, sw, se
        xx, yy = x0 + dx, y0 + dy
        ok = inside & (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        v = plane[:, np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].T                           # [M,C]
        out = out + np.where(ok, wgt, f32(0)).astype(f32)[:, None] * v
    return out.astype(f32)


def sample_from_planes(planes, coordinates, box_warp):
    """planes [N,3,C,H,W], coordinates [N,M,3] -> [N,3,M,C] (renderer.py:55-65)."""
    planes = np.asarray(planes, f32)
    coords = f32(2 / box_warp) * np.asarray(coordinates, f32)
    n, _, c, h, w = planes.shape
    m = coords.shape[1]
    out = np.empty((n, 3, m, c), f32)
    for b in range(n):
        for k, (a0, a1) in enumerate(PLANE_COORDS):
            out[b, k] = _bilinear_zeros(planes[b, k], coords[b, :, a0], coords[b, :, a1])
    return out


# ---------------------------------------------------------------------------------------------
# decoders: triplane.py:112-135, triplane_cond.py:859-970; FullyConnectedLayer networks_stylegan2.py:111-123
# ---------------------------------------------------------------------------------------------
def softplus(x):
    x = np.asarray(x, f32)
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, f32(20)), dtype=f32), dtype=f32)).astype(f32)


def sigmoid(x):
    x = np.asarray(x, f32)
    return (f32(1) / (f32(1) + np.exp(-x, dtype=f32))).astype(f32)


def mipnerf_sigmoid(x):
    return (sigmoid(x) * f32(1 + 2 * 0.001) - f32(0.001)).astype(f32)


def fc(x, weight, bias, lr_multiplier=1.0):
    """linear FullyConnectedLayer: x @ (W*gain).T + b*lr (addmm)."""
    w = np.asarray(weight, f32) * f32(lr_multiplier / np.sqrt(weight.shape[1]))
    b = np.asarray(bias, f32)
    if lr_multiplier != 1:
        b = b * f32(lr_multiplier)
    return (np.asarray(x, f32) @ w.T + b).astype(f32)


def decoder_forward(dec, sampled_features):
    """dec: dict(kind, nets=[{w1,b1,w2,b2}], lr_mul, sigmoid flags); sampled_features [N,3,M,C] -> rgb [N,M,Co], sigma [N,M,1]."""
    f = np.asarray(sampled_features, f32)
    x = ((f[:, 0] + f[:, 1]) + f[:, 2]) / f32(3)                                  # .mean(1)
    n, m, c = x.shape
    x = x.reshape(n * m, c)
    lr = dec.get('lr_mul', 1.0)
    outs = []
    for net in dec['nets']:
        h = softplus(fc(x, net['w1'], net['b1'], lr))
        outs.append(fc(h, net['w2'], net['b2'], lr).reshape(n, m, -1))
    kind = dec['kind']
    if kind == 'OSGDecoder':                                                      # triplane.py:123-135
        return mipnerf_sigmoid(outs[0][..., 1:]), outs[0][..., 0:1]
    if kind == 'OSGDecoder_semantic':                                             # triplane_cond.py:871-887
        rgb = mipnerf_sigmoid(outs[0][..., 1:]) if dec['sigmoid'] else outs[0][..., 1:]
        return rgb, outs[0][..., 0:1]
    if kind == 'OSGDecoder_semantic_lateSeparate':                                # triplane_cond.py:946-970
        rgb = mipnerf_sigmoid(outs[0][..., 1:])
        sem = mipnerf_sigmoid(outs[1][..., 1:]) if dec['sigmoid'] else outs[1][..., 1:]
        return np.concatenate([rgb, sem], -1).astype(f32), outs[1][..., 0:1]
    raise ValueError(kind)


def run_model(planes, dec, coords, box_warp):
    """renderer.py:142-148 without density noise."""
    return decoder_forward(dec, sample_from_planes(planes, coords, box_warp))


# ---------------------------------------------------------------------------------------------
# ray_marcher.py:25-57
# ---------------------------------------------------------------------------------------------
def ray_march(colors, densities, depths, white_back=False, clamp_range=None):
    """colors [B,R,S,C], densities [B,R,S,1], depths [B,R,S,1] -> rgb [B,R,C], depth [B,R,1], weights [B,R,S-1,1]."""
    colors, densities, depths = (np.asarray(a, f32) for a in (colors, densities, depths))
    deltas = depths[:, :, 1:] - depths[:, :, :-1]
    colors_mid = (colors[:, :, :-1] + colors[:, :, 1:]) / f32(2)
    dens_mid = (densities[:, :, :-1] + densities[:, :, 1:]) / f32(2)
    depths_mid = (depths[:, :, :-1] + depths[:, :, 1:]) / f32(2)
    dens_mid = softplus(dens_mid - f32(1))                                        # :33
    alpha = f32(1) - np.exp(-(dens_mid * deltas), dtype=f32)                      # :38-40
    shifted = np.concatenate([np.ones_like(alpha[:, :, :1]), f32(1) - alpha + f32(1e-10)], -2)
    trans = np.empty_like(shifted)                                                # cumprod, sequential (:42)
    acc = np.ones_like(shifted[:, :, 0])
    for i in range(shifted.shape[2]):
        acc = (acc * shifted[:, :, i]).astype(f32)
        trans[:, :, i] = acc
    weights = (alpha * trans[:, :, :-1]).astype(f32)
    rgb = np.zeros_like(colors_mid[:, :, 0])
    wsum = np.zeros_like(weights[:, :, 0])
    wd = np.zeros_like(weights[:, :, 0])
    for i in range(weights.shape[2]):                                             # sums, sequential (:44-46)
        rgb = rgb + weights[:, :, i] * colors_mid[:, :, i]
        wsum = wsum + weights[:, :, i]
        wd = wd + weights[:, :, i] * depths_mid[:, :, i]
    with np.errstate(divide='ignore', invalid='ignore'):
        depth = wd / wsum
    depth = np.where(np.isnan(depth), f32(np.inf), depth)                         # nan_to_num(nan=inf) (:49)
    lo, hi = (depths.min(), depths.max()) if clamp_range is None else clamp_range
    depth = np.clip(depth, lo, hi)                                                # :50 (also maps +-inf)
    if white_back:
        rgb = rgb + f32(1) - wsum                                                 # :53
    rgb = rgb * f32(2) - f32(1)                                                   # :55
    return rgb.astype(f32), depth.astype(f32), weights


# ---------------------------------------------------------------------------------------------
# renderer.py:194-253
# ---------------------------------------------------------------------------------------------
def sample_importance(z_vals, weights, u, eps=1e-5, return_debug=False):
    """z_vals [N,S], weights [N,S-1], u [N,Sf] (the torch.rand draw at :237) -> samples [N,Sf]."""
    z = np.asarray(z_vals, f32)
    w = np.asarray(weights, f32)
    u = np.ascontiguousarray(u, f32)
    n, s = z.shape
    ninf = f32(-np.inf)
    # max_pool1d(k=2, s=1, pad=1) -> S values; avg_pool1d(k=2, s=1) -> S-1 values; + 0.01   (:205-207)
    wp = np.maximum(np.concatenate([np.full((n, 1), ninf, f32), w], 1), np.concatenate([w, np.full((n, 1), ninf, f32)], 1))
    wa = (wp[:, :-1] + wp[:, 1:]) * f32(0.5) + f32(0.01)
    bins = f32(0.5) * (z[:, :-1] + z[:, 1:])                                      # :209
    om = wa[:, 1:-1] + f32(eps)                                                   # :210, :227
    k = om.shape[1]
    tot = np.zeros(n, f32)
    for i in range(k):                                                            # torch.sum, sequential
        tot = (tot + om[:, i]).astype(f32)
    pdf = (om / tot[:, None]).astype(f32)                                         # :228
    cdf = np.zeros((n, k + 1), f32)                                               # :229-230
    for i in range(k):
        cdf[:, i + 1] = (cdf[:, i] + pdf[:, i]).astype(f32)
    inds = np.empty(u.shape, np.int64)
    for r in range(n):
        inds[r] = np.searchsorted(cdf[r], u[r], side='right')                     # :240
    below = np.maximum(inds - 1, 0)
    above = np.minimum(inds, k)
    c0, c1 = np.take_along_axis(cdf, below, 1), np.take_along_axis(cdf, above, 1)
    b0, b1 = np.take_along_axis(bins, below, 1), np.take_along_axis(bins, above, 1)
    denom = c1 - c0
    denom = np.where(denom < f32(eps), f32(1), denom)                             # :248-249
    samples = (b0 + (u - c0) / denom * (b1 - b0)).astype(f32)                     # :252
    if return_debug:
        return samples, dict(inds=inds, cdf=cdf, bins=bins)
    return samples


def unify_samples(depths1, colors1, dens1, depths2, colors2, dens2):
    """cat + sort by depth + gather (renderer.py:157-167); stable order for ties."""
    d = np.concatenate([depths1, depths2], -2)
    c = np.concatenate([colors1, colors2], -2)
    s = np.concatenate([dens1, dens2], -2)
    idx = np.argsort(d, axis=-2, kind='stable')
    return (np.take_along_axis(d, idx, -2), np.take_along_axis(c, np.broadcast_to(idx, c.shape[:-1] + (1,)), -2),
            np.take_along_axis(s, idx, -2), idx[..., 0])


# ---------------------------------------------------------------------------------------------
# renderer.py:88-140
# ---------------------------------------------------------------------------------------------
def importance_renderer(planes, dec, ray_origins, ray_directions, depths_coarse, u, opts, return_debug=False):
    """ImportanceRenderer.forward with the two random draws supplied explicitly.
    planes [B,3,C,H,W]; rays [B,R,3]; depths_coarse [B,R,Sc,1]; u [B*R,Sf] or None."""
    planes = np.asarray(planes, f32)
    o = np.asarray(ray_origins, f32)
    d = np.asarray(ray_directions, f32)
    dc = np.asarray(depths_coarse, f32)
    b, r, sc, _ = dc.shape
    bw = opts['box_warp']
    wb = bool(opts.get('white_back', False))

    def shade(depths, count):
        coords = (o[:, :, None, :] + depths * d[:, :, None, :]).reshape(b, -1, 3)
        rgb, sigma = run_model(planes, dec, coords, bw)
        return rgb.reshape(b, r, count, -1), sigma.reshape(b, r, count, 1)

    colors_c, dens_c = shade(dc, sc)
    dbg = {}
    sf = 0 if u is None else u.shape[-1]
    if sf > 0:
        _, _, w_c = ray_march(colors_c, dens_c, dc, wb)
        dfine, idbg = sample_importance(dc.reshape(b * r, sc), w_c.reshape(b * r, -1), u, return_debug=True)
        dfine = dfine.reshape(b, r, sf, 1)
        colors_f, dens_f = shade(dfine, sf)
        all_d, all_c, all_s, perm = unify_samples(dc, colors_c, dens_c, dfine, colors_f, dens_f)
        rgb, depth, w = ray_march(all_c, all_s, all_d, wb)
        dbg.update(weights_coarse=w_c[..., 0], depths_fine=dfine[..., 0], inds=idbg['inds'].reshape(b, r, sf), perm=perm,
                   dens_coarse=dens_c[..., 0], dens_fine=dens_f[..., 0])
    else:
        rgb, depth, w = ray_march(colors_c, dens_c, dc, wb)
        dbg.update(dens_coarse=dens_c[..., 0])
    dbg['weights_final'] = w[..., 0]
    wsum = np.zeros_like(w[:, :, 0])
    for i in range(w.shape[2]):                                                   # weights.sum(2) (:140)
        wsum = wsum + w[:, :, i]
    if return_debug:
        return rgb, depth, wsum.astype(f32), dbg
    return rgb, depth, wsum.astype(f32)


# ---------------------------------------------------------------------------------------------
# renderer.py:256-337  ImportanceSemanticRenderer (two plane sets, two decoders)
# ---------------------------------------------------------------------------------------------
def run_model_semantic(planes_texture, planes_semantic, dec_texture, dec_semantic, coords, box_warp):
    """renderer.py:324-337: sigma and semantics from the semantic decoder (semantic planes), colour from the texture
    decoder on cat(texture, semantic) features. Returns (rgb, sigma, semantic)."""
    f_tex = sample_from_planes(planes_texture, coords, box_warp)
    f_sem = sample_from_planes(planes_semantic, coords, box_warp)
    sem, sigma = decoder_forward(dec_semantic, f_sem)
    rgb, _ = decoder_forward(dec_texture, np.concatenate([f_tex, f_sem], -1))
    return rgb, sigma, sem


def importance_semantic_renderer(planes_texture, planes_semantic, dec_texture, dec_semantic, ray_origins, ray_directions,
                                 depths_coarse, u, opts):
    """ImportanceSemanticRenderer.forward (:262-322) with the two random draws supplied explicitly; the coarse weights
    come from the colours (:296), the composited feature is cat(colour, semantic) (:292, :314)."""
    pt, ps = np.asarray(planes_texture, f32), np.asarray(planes_semantic, f32)
    o = np.asarray(ray_origins, f32)
    d = np.asarray(ray_directions, f32)
    dc = np.asarray(depths_coarse, f32)
    b, r, sc, _ = dc.shape
    bw = opts['box_warp']
    wb = bool(opts.get('white_back', False))

    def shade(depths, count):
        coords = (o[:, :, None, :] + depths * d[:, :, None, :]).reshape(b, -1, 3)
        rgb, sigma, sem = run_model_semantic(pt, ps, dec_texture, dec_semantic, coords, bw)
        rgb = rgb.reshape(b, r, count, -1)
        return rgb, sigma.reshape(b, r, count, 1), np.concatenate([rgb, sem.reshape(b, r, count, -1)], -1)

    colors_c, dens_c, feats_c = shade(dc, sc)
    sf = 0 if u is None else u.shape[-1]
    if sf > 0:
        _, _, w_c = ray_march(colors_c, dens_c, dc, wb)
        dfine = sample_importance(dc.reshape(b * r, sc), w_c.reshape(b * r, -1), u).reshape(b, r, sf, 1)
        _, dens_f, feats_f = shade(dfine, sf)
        all_d, all_f, all_s, _ = unify_samples(dc, feats_c, dens_c, dfine, feats_f, dens_f)
        feat, depth, w = ray_march(all_f, all_s, all_d, wb)
    else:
        feat, depth, w = ray_march(feats_c, dens_c, dc, wb)
    wsum = np.zeros_like(w[:, :, 0])
    for i in range(w.shape[2]):
        wsum = wsum + w[:, :, i]
    return feat, depth, wsum.astype(f32)


# ---------------------------------------------------------------------------------------------
# First-order gradients of the two differentiable renderer stages (what autograd derives from renderer.py:55-65 and
# ray_marcher.py:25-57); importance sampling is under no_grad / detached in the reference (renderer.py:198,211).
# ---------------------------------------------------------------------------------------------
def sample_from_planes_backward(grad_out, plane_shape, coordinates, box_warp):
    """grad_out [N,3,M,C] -> gradient w.r.t. planes [N,3,C,H,W]: the bilinear taps scattered back (zero padding)."""
    g = np.asarray(grad_out, np.float64)
    n, _, c, h, w = plane_shape
    coords = f32(2 / box_warp) * np.asarray(coordinates, f32)
    out = np.zeros(plane_shape, np.float64)
    for b in range(n):
        for k, (a0, a1) in enumerate(PLANE_COORDS):
            gx, gy = coords[b, :, a0], coords[b, :, a1]
            ix = ((gx + f32(1)) * f32(w) - f32(1)) * f32(0.5)
            iy = ((gy + f32(1)) * f32(h) - f32(1)) * f32(0.5)
            x0f, y0f = np.floor(ix), np.floor(iy)
            ax, bx = (x0f + f32(1)) - ix, ix - x0f
            ay, by = (y0f + f32(1)) - iy, iy - y0f
            inside = (ix > -1) & (ix < w) & (iy > -1) & (iy < h)
            x0 = np.where(inside, x0f, 0).astype(np.int64)
            y0 = np.where(inside, y0f, 0).astype(np.int64)
            for dy, dx, wgt in ((0, 0, ax * ay), (0, 1, bx * ay), (1, 0, ax * by), (1, 1, bx * by)):
                xx, yy = x0 + dx, y0 + dy
                ok = inside & (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
                wv = np.where(ok, wgt, f32(0)).astype(np.float64)
                np.add.at(out[b, k], (slice(None), np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)), (wv[:, None] * g[b, k]).T)
    return out.astype(f32)


def decoder_backward(dec, sampled_features, g_rgb, g_sigma):
    """First-order gradients of `decoder_forward` (what autograd derives from triplane.py:122-135 / triplane_cond.py:869-970 with
    FullyConnectedLayer networks_stylegan2.py:111-123 and torch.nn.Softplus): sampled_features [N,3,M,32], upstream g_rgb [N,M,Co],
    g_sigma [N,M,1] -> (g_features [N,3,M,32], [per net dict(w1, b1, w2, b2)] = gradients of the RAW parameters). float64."""
    f = np.asarray(sampled_features, np.float64)
    x = ((f[:, 0] + f[:, 1]) + f[:, 2]) / 3.0
    n, m, c = x.shape
    x = x.reshape(n * m, c)
    lr = float(dec.get('lr_mul', 1.0))
    kind = dec['kind']
    g_rgb = np.asarray(g_rgb, np.float64).reshape(n * m, -1)
    g_sig = np.asarray(g_sigma, np.float64).reshape(n * m, 1)
    gx = np.zeros_like(x)
    grads = []
    for k, net in enumerate(dec['nets']):
        w1, b1, w2, b2 = (np.asarray(net[q], np.float64) for q in ('w1', 'b1', 'w2', 'b2'))
        g1, g2 = lr / np.sqrt(w1.shape[1]), lr / np.sqrt(w2.shape[1])                # weight_gain; bias_gain = lr (:108-109)
        a1 = x @ (w1 * g1).T + b1 * lr
        z = np.exp(np.minimum(a1, 20.0))
        h = np.where(a1 > 20, a1, np.log1p(z))                                       # Softplus(beta 1, threshold 20)
        dsp = np.where(a1 > 20, 1.0, z / (z + 1.0))
        o = h @ (w2 * g2).T + b2 * lr
        # upstream gradient of this net's 33 outputs
        go = np.zeros_like(o)
        if kind == 'OSGDecoder' or (kind == 'OSGDecoder_semantic'):
            sig_on, grgb, has_sigma = (True if kind == 'OSGDecoder' else bool(dec['sigmoid'])), g_rgb, True
        elif kind == 'OSGDecoder_semantic_lateSeparate':                             # net 0: colours, net 1: semantics + sigma
            sig_on = True if k == 0 else bool(dec['sigmoid'])
            grgb, has_sigma = g_rgb[:, 32 * k:32 * (k + 1)], k == 1
        else:
            raise ValueError(kind)
        if sig_on:
            sg = 1.0 / (1.0 + np.exp(-o[:, 1:]))
            go[:, 1:] = grgb * (1 + 2 * 0.001) * sg * (1.0 - sg)                     # y = sigmoid(o) * 1.002 - 0.001
        else:
            go[:, 1:] = grgb
        if has_sigma:
            go[:, 0:1] = g_sig
        gh = go @ (w2 * g2)
        ga = gh * dsp
        gx += ga @ (w1 * g1)
        grads.append(dict(w1=(ga.T @ x) * g1, b1=ga.sum(0) * lr, w2=(go.T @ h) * g2, b2=go.sum(0) * lr))
    gf = np.broadcast_to((gx / 3.0).reshape(n, 1, m, c), (n, 3, m, c)).copy()
    return gf, grads


def ray_march_backward(colors, densities, depths, g_rgb, g_depth, g_weights, white_back=False, clamp_range=None):
    """Gradients of `ray_march` w.r.t. colors [B,R,S,C] and densities [B,R,S,1] for upstream gradients of its three
    outputs (g_rgb [B,R,C], g_depth [B,R,1] or None, g_weights [B,R,S-1,1] or None). Depths carry no gradient in the
    reference's pipeline. The depth term follows autograd: it passes only where the composite depth is finite and inside
    the clamp range, and is skipped for rays whose upstream depth gradient is exactly zero."""
    c = np.asarray(colors, np.float64)
    s = np.asarray(densities, np.float64)[..., 0]
    z = np.asarray(depths, np.float64)[..., 0]
    delta = z[:, :, 1:] - z[:, :, :-1]
    cmid = (c[:, :, :-1] + c[:, :, 1:]) / 2
    smid = (s[:, :, :-1] + s[:, :, 1:]) / 2
    zmid = (z[:, :, :-1] + z[:, :, 1:]) / 2
    x = smid - 1
    sp = np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))
    e = np.exp(-sp * delta)
    alpha = 1 - e
    a = 1 - alpha + 1e-10
    trans = np.concatenate([np.ones_like(a[:, :, :1]), np.cumprod(a, -1)[:, :, :-1]], -1)
    w = alpha * trans
    wsum = w.sum(-1)
    G = 2 * np.asarray(g_rgb, np.float64)
    gw = np.einsum('brc,bric->bri', G, cmid)
    if white_back:
        gw = gw - G.sum(-1)[..., None]
    if g_weights is not None:
        gw = gw + np.asarray(g_weights, np.float64)[..., 0]
    if g_depth is not None:
        gd = np.asarray(g_depth, np.float64)[..., 0]
        with np.errstate(divide='ignore', invalid='ignore'):
            draw = (w * zmid).sum(-1) / wsum
        lo, hi = (np.asarray(depths, f32).min(), np.asarray(depths, f32).max()) if clamp_range is None else clamp_range
        ok = np.isfinite(draw) & (draw >= lo) & (draw <= hi) & (gd != 0)
        with np.errstate(divide='ignore', invalid='ignore'):
            term = gd[..., None] * (zmid - draw[..., None]) / wsum[..., None]
        gw = gw + np.where(ok[..., None], term, 0)
    g_cmid = w[..., None] * G[:, :, None, :]
    gww = gw * w
    suffix = np.concatenate([np.cumsum(gww[:, :, ::-1], -1)[:, :, ::-1][:, :, 1:], np.zeros_like(gww[:, :, :1])], -1)
    g_alpha = gw * trans - suffix / a
    g_sp = g_alpha * e * delta
    g_smid = g_sp * np.where(x > 20, 1.0, 1 / (1 + np.exp(-np.clip(x, -80, 80))))
    g_c = np.zeros_like(c)
    g_c[:, :, :-1] += g_cmid / 2
    g_c[:, :, 1:] += g_cmid / 2
    g_s = np.zeros_like(s)
    g_s[:, :, :-1] += g_smid / 2
    g_s[:, :, 1:] += g_smid / 2
    return g_c.astype(f32), g_s[..., None].astype(f32)
