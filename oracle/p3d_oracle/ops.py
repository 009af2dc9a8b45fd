"""numpy restatement of the reference's custom ops: torch_utils/ops/{bias_act,upfirdn2d,conv2d_resample,fma}.py
and modulated_conv2d (training/networks_stylegan2.py:34-91). float32 unless the input is float64."""
import os

import numpy as np

f32 = np.float32

# ---------------------------------------------------------------------------------------------
# bias_act.py:23-33, :93-122 (forward) and bias_act.cu:44-146 (gradient forms)
# ---------------------------------------------------------------------------------------------
_SQ2 = float(np.sqrt(2))
ACT = {  # name: (def_alpha, def_gain, cuda_idx, ref, has_2nd_grad)
    'linear': (0, 1, 1, '', False), 'relu': (0, _SQ2, 2, 'y', False), 'lrelu': (0.2, _SQ2, 3, 'y', False),
    'tanh': (0, 1, 4, 'y', True), 'sigmoid': (0, 1, 5, 'y', True), 'elu': (0, 1, 6, 'y', True),
    'selu': (0, 1, 7, 'y', True), 'softplus': (0, 1, 8, 'y', True), 'swish': (0, _SQ2, 9, 'x', True),
}
_SELU_L = 1.0507009873554804934193349852946
_SELU_A = 1.6732632423543772848170429916717


def _act_fwd(name, x, alpha):
    t = x.dtype.type
    if name == 'linear':
        return x
    if name == 'relu':
        return np.maximum(x, t(0))
    if name == 'lrelu':
        return np.where(x > 0, x, x * t(alpha))
    if name == 'tanh':
        return np.tanh(x)
    if name == 'sigmoid':
        return t(1) / (t(1) + np.exp(-x))
    if name == 'elu':
        return np.where(x >= 0, x, np.expm1(np.minimum(x, t(0))))
    if name == 'selu':
        return t(_SELU_L) * np.where(x >= 0, x, t(_SELU_A) * np.expm1(np.minimum(x, t(0))))
    if name == 'softplus':
        return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, t(20)))))
    if name == 'swish':
        return x / (t(1) + np.exp(-x))
    raise ValueError(name)


def _bshape(x, dim):
    s = [1] * x.ndim
    s[dim] = -1
    return s


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """_bias_act_ref (bias_act.py:93-122)."""
    da, dg = ACT[act][0], ACT[act][1]
    alpha = float(da if alpha is None else alpha)
    gain = float(dg if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    x = np.asarray(x)
    t = x.dtype.type
    cdt = np.float64 if x.dtype == np.float64 else np.float32
    y = x.astype(cdt)
    if b is not None:
        y = y + np.asarray(b).astype(cdt).reshape(_bshape(x, dim))
    y = _act_fwd(act, y, alpha)
    if gain != 1:
        y = y * cdt(gain)
    if clamp >= 0:
        y = np.clip(y, cdt(-clamp), cdt(clamp))
    return y.astype(t)


def bias_act_grad(dy, x, b, y, dim=1, act='linear', alpha=None, gain=None, clamp=None, order=1, dy1=None, plugin_semantics=False):
    """First (order=1: dx given dy) and second (order=2: d_x given d_dx=`dy` and the first-order `dy1`) gradient
    forms of the plugin (bias_act.cu:44-146 with grad = 1 / 2). x: forward input, y: forward output.
    plugin_semantics: the CUDA autograd wrapper only saves y when the activation's `ref` contains 'y' (bias_act.py:160-163),
    so for 'linear' the plugin sees yref = 0 and the clamp does NOT mask the gradient -- unlike autograd through `_ref`."""
    da, dg = ACT[act][0], ACT[act][1]
    alpha = float(da if alpha is None else alpha)
    gain = float(dg if gain is None else gain)
    clamp = float(-1 if clamp is None else clamp)
    g = np.asarray(dy)
    t = g.dtype.type
    cdt = np.float64 if g.dtype == np.float64 else np.float32
    g = g.astype(cdt)
    one = cdt(1)
    xr = None if x is None else np.asarray(x).astype(cdt)
    if xr is not None and b is not None:
        xr = xr + np.asarray(b).astype(cdt).reshape(_bshape(g, dim))
    yr = None if y is None else np.asarray(y).astype(cdt)
    if plugin_semantics and 'y' not in ACT[act][3]:
        yr = np.zeros_like(g)
    yy = (yr / cdt(gain)) if (yr is not None and gain != 0) else (np.zeros_like(g))
    if act == 'linear':
        r = g if order == 1 else np.zeros_like(g)
    elif act == 'relu':
        r = np.where(yy > 0, g, 0) if order == 1 else np.zeros_like(g)
    elif act == 'lrelu':
        r = np.where(yy > 0, g, g * cdt(alpha)) if order == 1 else np.zeros_like(g)
    elif act == 'tanh':
        r = g * (one - yy * yy) * (one if order == 1 else (-2 * yy))
    elif act == 'sigmoid':
        r = g * yy * (one - yy) * (one if order == 1 else (one - 2 * yy))
    elif act == 'elu':
        r = np.where(yy >= 0, g if order == 1 else 0, g * (yy + one))
    elif act == 'selu':
        la = cdt(_SELU_L * _SELU_A)
        r = np.where(yy >= 0, g * cdt(_SELU_L) if order == 1 else 0, g * (yy + la))
    elif act == 'softplus':
        c = np.exp(-yy)
        r = g * (one - c) if order == 1 else g * c * (one - c)
    elif act == 'swish':
        c = np.exp(xr)
        d = c + one
        if order == 1:
            r = np.where(xr > 40, g, g * c * (xr + d) / (d * d))
        else:
            r = np.where(xr > 40, 0, g * c * (xr * (2 - d) + 2 * d) / (d * d * d))
        yr = np.where(xr < -80, 0, xr / (np.exp(-xr) + one) * cdt(gain))
    else:
        raise ValueError(act)
    r = r * cdt(gain)
    if order == 2:
        r = r * np.asarray(dy1).astype(cdt)
    if clamp >= 0:
        r = np.where((yr > -clamp) & (yr < clamp), r, 0)
    return r.astype(t)


# ---------------------------------------------------------------------------------------------
# upfirdn2d.py:72-115 (setup_filter), :169-213 (_upfirdn2d_ref), :279-389 (wrappers)
# ---------------------------------------------------------------------------------------------
def setup_filter(f, normalize=True, flip_filter=False, gain=1, separable=None):
    f = np.asarray(1 if f is None else f, f32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = (f.ndim == 1 and f.size >= 8)
    if f.ndim == 1 and not separable:
        f = np.outer(f, f).astype(f32)
    if normalize:
        f = (f / f.sum()).astype(f32)
    if flip_filter:
        f = np.flip(f, tuple(range(f.ndim)))
    return (f * f32(gain ** (f.ndim / 2))).astype(f32)


def _pad4(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    if len(padding) == 2:
        padding = [padding[0], padding[0], padding[1], padding[1]]
    return [int(p) for p in padding]


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    x = np.asarray(x)
    t = x.dtype.type
    cdt = np.float64 if x.dtype == np.float64 else np.float32
    n, c, ih, iw = x.shape
    f = np.ones((1, 1), f32) if f is None else np.asarray(f, f32)
    upx, upy = _pair(up)
    downx, downy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    z = np.zeros((n, c, ih * upy, iw * upx), cdt)
    z[:, :, ::upy, ::upx] = x
    z = np.pad(z, ((0, 0), (0, 0), (max(py0, 0), max(py1, 0)), (max(px0, 0), max(px1, 0))))
    z = z[:, :, max(-py0, 0): z.shape[2] - max(-py1, 0), max(-px0, 0): z.shape[3] - max(-px1, 0)]
    k = (f * f32(gain ** (f.ndim / 2))).astype(cdt)
    if not flip_filter:
        k = np.flip(k, tuple(range(k.ndim)))

    torch = _aten()
    if torch is not None:
        # the reference filters with a depthwise ATen convolution (upfirdn2d.py:203-209)
        zt = torch.from_numpy(np.ascontiguousarray(z))
        kt = torch.from_numpy(k.copy())
        if k.ndim == 2:
            zt = torch.nn.functional.conv2d(zt, kt[None, None].repeat(c, 1, 1, 1), groups=c)
        else:
            zt = torch.nn.functional.conv2d(zt, kt[None, None, None, :].repeat(c, 1, 1, 1), groups=c)
            zt = torch.nn.functional.conv2d(zt, kt[None, None, :, None].repeat(c, 1, 1, 1), groups=c)
        return zt.numpy()[:, :, ::downy, ::downx].astype(t)
    if k.ndim == 2:
        kh, kw = k.shape
        oh, ow = z.shape[2] - kh + 1, z.shape[3] - kw + 1
        out = np.zeros((n, c, oh, ow), cdt)
        for j in range(kh):
            for i in range(kw):
                out += k[j, i] * z[:, :, j:j + oh, i:i + ow]
    else:
        kk = k.shape[0]
        ow = z.shape[3] - kk + 1
        tmp = np.zeros(z.shape[:3] + (ow,), cdt)
        for i in range(kk):
            tmp += k[i] * z[:, :, :, i:i + ow]
        oh = z.shape[2] - kk + 1
        out = np.zeros((n, c, oh, ow), cdt)
        for j in range(kk):
            out += k[j] * tmp[:, :, j:j + oh, :]
    return out[:, :, ::downy, ::downx].astype(t)


def _fsize(f):
    if f is None:
        return 1, 1
    f = np.asarray(f)
    return int(f.shape[-1]), int(f.shape[0])


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1):
    upx, upy = _pair(up)
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = _fsize(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1):
    dx, dy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = _fsize(f)
    p = [px0 + (fw - dx + 1) // 2, px1 + (fw - dx) // 2, py0 + (fh - dy + 1) // 2, py1 + (fh - dy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


def filter2d(x, f, padding=0, flip_filter=False, gain=1):
    px0, px1, py0, py1 = _pad4(padding)
    fw, fh = _fsize(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


# ---------------------------------------------------------------------------------------------
# dense convolutions (F.conv2d / F.conv_transpose2d semantics), im2col + matmul
# ---------------------------------------------------------------------------------------------
def _aten():
    """ATen's CPU convolution -- the very call the reference makes (torch_utils/ops/conv2d_gradfix.py:40-45).
    Default backend of the two dense convolutions below (it keeps the CPU baseline representative of the
    reference's CPU speed); P3D_ORACLE_NUMPY_CONV=1 selects the explicit im2col + matmul restatement instead.
    tests/test_oracle_golden.py checks that the two agree."""
    if os.environ.get('P3D_ORACLE_NUMPY_CONV', '0') == '1':
        return None
    try:
        import torch
        return torch
    except ImportError:
        return None


def conv2d(x, w, stride=1, padding=0, groups=1):
    """Cross-correlation. x [N,Cin,H,W], w [Cout,Cin/groups,kh,kw]."""
    x = np.asarray(x)
    torch = _aten()
    if torch is not None:
        pad = tuple(padding) if not isinstance(padding, int) else padding
        y = torch.nn.functional.conv2d(torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(np.ascontiguousarray(w, x.dtype)),
                                       stride=stride, padding=pad, groups=groups)
        return y.numpy()
    t = x.dtype.type
    cdt = np.float64 if x.dtype == np.float64 else np.float32
    x = x.astype(cdt)
    w = np.asarray(w).astype(cdt)
    n, cin, h, wd = x.shape
    cout, cin_g, kh, kw = w.shape
    ph, pw = _pair(padding) if not isinstance(padding, int) else (padding, padding)
    xp = np.pad(x, ((0, 0), (0, 0), (ph, ph), (pw, pw)))
    oh = (h + 2 * ph - kh) // stride + 1
    ow = (wd + 2 * pw - kw) // stride + 1
    out = np.empty((n, cout, oh, ow), cdt)
    cog = cout // groups
    for g in range(groups):
        xs = xp[:, g * cin_g:(g + 1) * cin_g]
        cols = np.empty((n, cin_g, kh, kw, oh, ow), cdt)
        for j in range(kh):
            for i in range(kw):
                cols[:, :, j, i] = xs[:, :, j:j + stride * oh:stride, i:i + stride * ow:stride]
        wm = w[g * cog:(g + 1) * cog].reshape(cog, -1)
        out[:, g * cog:(g + 1) * cog] = np.einsum('ok,nkp->nop', wm, cols.reshape(n, cin_g * kh * kw, oh * ow),
                                                  optimize=True).reshape(n, cog, oh, ow)
    return out.astype(t)


def conv_transpose2d(x, w, stride=1, padding=0, groups=1):
    """x [N,Cin,H,W], w [Cin,Cout/groups,kh,kw] (PyTorch layout); output (H-1)*stride - 2*pad + k."""
    x = np.asarray(x)
    torch = _aten()
    if torch is not None:
        pad = tuple(padding) if not isinstance(padding, int) else padding
        y = torch.nn.functional.conv_transpose2d(torch.from_numpy(np.ascontiguousarray(x)),
                                                 torch.from_numpy(np.ascontiguousarray(w, x.dtype)), stride=stride, padding=pad,
                                                 groups=groups)
        return y.numpy()
    t = x.dtype.type
    cdt = np.float64 if x.dtype == np.float64 else np.float32
    x = x.astype(cdt)
    w = np.asarray(w).astype(cdt)
    n, cin, h, wd = x.shape
    _, cog, kh, kw = w.shape
    ph, pw = _pair(padding) if not isinstance(padding, int) else (padding, padding)
    cin_g = cin // groups
    fh, fw = (h - 1) * stride + kh, (wd - 1) * stride + kw
    full = np.zeros((n, cog * groups, fh, fw), cdt)
    for g in range(groups):
        xs = x[:, g * cin_g:(g + 1) * cin_g].reshape(n, cin_g, h * wd)
        wm = w[g * cin_g:(g + 1) * cin_g].reshape(cin_g, cog * kh * kw)
        contrib = np.einsum('ck,ncp->nkp', wm, xs, optimize=True).reshape(n, cog, kh, kw, h, wd)
        for j in range(kh):
            for i in range(kw):
                full[:, g * cog:(g + 1) * cog, j:j + stride * h:stride, i:i + stride * wd:stride] += contrib[:, :, j, i]
    return full[:, :, ph:fh - ph, pw:fw - pw].astype(t)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """conv2d_resample.py:48-143."""
    w = np.asarray(w)
    cout, cin_g, kh, kw = w.shape
    fw, fh = _fsize(f)
    px0, px1, py0, py1 = _pad4(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2; py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2; py0 += (fh - down + 1) // 2; py1 += (fh - down) // 2

    def conv(xx, ww, stride=1, pad=(0, 0), transpose=False, flip=True):
        if not flip and (ww.shape[2] > 1 or ww.shape[3] > 1):
            ww = ww[:, :, ::-1, ::-1]
        return (conv_transpose2d if transpose else conv2d)(xx, ww, stride=stride, padding=pad, groups=groups)

    if kh == 1 and kw == 1 and down > 1 and up == 1:
        x = upfirdn2d(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return conv(x, w, flip=flip_weight)
    if kh == 1 and kw == 1 and up > 1 and down == 1:
        x = conv(x, w, flip=flip_weight)
        return upfirdn2d(x, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:
        x = upfirdn2d(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return conv(x, w, stride=down, flip=flip_weight)
    if up > 1:
        if groups == 1:
            wt = w.transpose(1, 0, 2, 3)
        else:
            wt = w.reshape(groups, cout // groups, cin_g, kh, kw).transpose(0, 2, 1, 3, 4).reshape(groups * cin_g, cout // groups, kh, kw)
        px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = conv(x, wt, stride=up, pad=(pyt, pxt), transpose=True, flip=(not flip_weight))
        x = upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
        return x
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return conv(x, w, pad=(py0, px0), flip=flip_weight)
    x = upfirdn2d(x, (f if up > 1 else None), up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = conv(x, w, flip=flip_weight)
    if down > 1:
        x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
    return x


# ---------------------------------------------------------------------------------------------
# networks_stylegan2.py:34-91
# ---------------------------------------------------------------------------------------------
def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    x = np.asarray(x, f32)
    weight = np.asarray(weight, f32)
    styles = np.asarray(styles, f32)
    n = x.shape[0]
    cout, cin, kh, kw = weight.shape
    w = dcoefs = None
    if demodulate or fused_modconv:
        w = weight[None] * styles.reshape(n, 1, -1, 1, 1)
    if demodulate:
        dcoefs = (f32(1) / np.sqrt((w * w).sum(axis=(2, 3, 4), dtype=f32) + f32(1e-8))).astype(f32)
    if demodulate and fused_modconv:
        w = w * dcoefs.reshape(n, -1, 1, 1, 1)
    if not fused_modconv:
        x = x * styles.reshape(n, -1, 1, 1)
        x = conv2d_resample(x, weight, f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = x * dcoefs.reshape(n, -1, 1, 1) + np.asarray(noise, f32)
        elif demodulate:
            x = x * dcoefs.reshape(n, -1, 1, 1)
        elif noise is not None:
            x = x + np.asarray(noise, f32)
        return x.astype(f32)
    xs = x.reshape(1, -1, *x.shape[2:])
    ws = w.reshape(-1, cin, kh, kw).astype(f32)
    y = conv2d_resample(xs, ws, f=resample_filter, up=up, down=down, padding=padding, groups=n, flip_weight=flip_weight)
    y = y.reshape(n, -1, *y.shape[2:])
    if noise is not None:
        y = y + np.asarray(noise, f32)
    return y.astype(f32)


# ---------------------------------------------------------------------------------------------
# filtered_lrelu.py:123-155 (generic composition)
# ---------------------------------------------------------------------------------------------
def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=_SQ2, slope=0.2, clamp=None, flip_filter=False):
    px0, px1, py0, py1 = _pad4(padding)
    x = bias_act(x, b)
    x = upfirdn2d(x, fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = bias_act(x, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
    return upfirdn2d(x, fd, down=down, flip_filter=flip_filter)


# ---------------------------------------------------------------------------------------------
# training/loss_utils.py:4-18 (cross_entropy2d, same-size case) = F.cross_entropy(reduction='mean') per pixel
# ---------------------------------------------------------------------------------------------
def cross_entropy2d(logits, target, weight=None, ignore_index=-100):
    """logits [N,C,H,W], target [N,H,W] int -> (loss, grad_logits). loss = sum_i w[t_i] (lse_i - x_i[t_i]) / sum_i w[t_i]
    over the pixels with t_i != ignore_index; grad = w[t] (softmax - onehot) / sum w."""
    x = np.asarray(logits, np.float64)
    t = np.asarray(target, np.int64)
    n, c, h, w = x.shape
    m = x.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(x - m).sum(axis=1))
    valid = t != ignore_index
    tc = np.where(valid, t, 0)
    xt = np.take_along_axis(x, tc[:, None], axis=1)[:, 0]
    wt = np.ones_like(lse) if weight is None else np.asarray(weight, np.float64)[tc]
    wt = wt * valid
    denom = wt.sum()
    loss = (wt * (lse - xt)).sum() / denom
    sm = np.exp(x - lse[:, None])
    onehot = (np.arange(c)[None, :, None, None] == tc[:, None]).astype(np.float64)
    grad = wt[:, None] * (sm - onehot) / denom
    return np.float32(loss), grad.astype(np.float32)
