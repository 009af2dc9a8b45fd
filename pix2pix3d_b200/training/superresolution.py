"""Super-resolution stacks that upsample the neural-rendered feature image.

Mirror of the reference's training/superresolution.py: every variant is two StyleGAN2 synthesis blocks driven
by the last w (`ws[:, -1:]` repeated three times), preceded by an antialiased bilinear resize when the input
is not at the expected resolution (:29-354). Class names, constructor arguments and module names match.
"""
import torch

from ..torch_utils import persistence
from ..torch_utils.ops import upfirdn2d
from ..torch_utils.ops.resize import interpolate_bilinear
from .networks_stylegan2 import SynthesisBlock, _block_forward, _block_setup


@persistence.persistent_class
class SynthesisBlockNoUp(torch.nn.Module):
    """SynthesisBlock whose conv0 and skip path keep the resolution (reference :191-289)."""

    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip',
                 resample_filter=[1, 3, 3, 1], conv_clamp=256, use_fp16=False, fp16_channels_last=False,
                 fused_modconv_default=True, **layer_kwargs):
        super().__init__()
        _block_setup(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture,
                     resample_filter, conv_clamp, use_fp16, fp16_channels_last, fused_modconv_default, False, layer_kwargs)

    def forward(self, x, img, ws, force_fp32=False, fused_modconv=None, update_emas=False, **layer_kwargs):
        _ = update_emas
        return _block_forward(self, x, img, ws, force_fp32, fused_modconv, False, layer_kwargs)

    def extra_repr(self):
        return f'resolution={self.resolution:d}, architecture={self.architecture:s}'


class _TwoBlockSR(torch.nn.Module):
    """Common body: resize rule + block0 + block1."""

    # subclasses set these
    OUT_RES = None          # required img_resolution
    IN_RES = None           # resolution block0 expects
    C0 = C1 = None          # block output channels
    BLOCK0_UP = True        # block0 doubles the resolution?
    RESIZE_ONLY_IF_SMALLER = False
    HAS_FILTER_BUFFER = False

    def _build(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, out_channels, block_kwargs):
        assert img_resolution == self.OUT_RES
        use_fp16 = sr_num_fp16_res > 0
        clamp = 256 if use_fp16 else None
        self.input_resolution = self.IN_RES
        self.sr_antialias = sr_antialias
        res0 = self.IN_RES * 2 if self.BLOCK0_UP else self.IN_RES
        block0_cls = SynthesisBlock if self.BLOCK0_UP else SynthesisBlockNoUp
        self.block0 = block0_cls(channels, self.C0, w_dim=512, resolution=res0, img_channels=out_channels, is_last=False,
                                 use_fp16=use_fp16, conv_clamp=clamp, **block_kwargs)
        self.block1 = SynthesisBlock(self.C0, self.C1, w_dim=512, resolution=res0 * 2, img_channels=out_channels,
                                     is_last=True, use_fp16=use_fp16, conv_clamp=clamp, **block_kwargs)
        if self.HAS_FILTER_BUFFER:
            self.register_buffer('resample_filter', upfirdn2d.setup_filter([1, 3, 3, 1]))

    def forward(self, rgb, x, ws, **block_kwargs):
        ws = ws[:, -1:, :].repeat(1, 3, 1)
        need = (x.shape[-1] < self.input_resolution) if self.RESIZE_ONLY_IF_SMALLER else (x.shape[-1] != self.input_resolution)
        if need:
            size = (self.input_resolution, self.input_resolution)
            x = interpolate_bilinear(x, size, antialias=self.sr_antialias)
            rgb = interpolate_bilinear(rgb, size, antialias=self.sr_antialias)
        if ws.device.type == 'cuda':
            from .. import engine, tcconv
            noise_mode = block_kwargs.get('noise_mode', 'random')
            extra = set(block_kwargs) - {'noise_mode', 'force_fp32', 'fused_modconv', 'update_emas'}
            if (not extra and not engine.grad_needed(self, ws, x, rgb) and engine.block_supported(self.block0, ws, noise_mode, False)
                    and engine.block_supported(self.block1, ws, noise_mode, False)):
                img, img0 = engine.superresolution(self, rgb.float().permute(0, 2, 3, 1).contiguous(), x.float(), ws, noise_mode=noise_mode,
                                                   force_fp32=bool(block_kwargs.get('force_fp32', False)), return_block0_image=True)
                if not self.BLOCK0_UP and not need:
                    # reference quirk: SynthesisBlockNoUp does `img.add_(y)` on the tensor it was handed (superresolution.py:283),
                    # i.e. on the caller's image_raw / semantic_raw view; keep the caller's tensor consistent with that
                    rgb.copy_(img0.permute(0, 3, 1, 2))
                return tcconv.nhwc_to_nchw_f32(img)
        x, rgb = self.block0(x, rgb, ws, **block_kwargs)
        x, rgb = self.block1(x, rgb, ws, **block_kwargs)
        return rgb


def _rgb_ctor(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, num_fp16_res=4, conv_clamp=None,
              channel_base=None, channel_max=None, **block_kwargs):
    torch.nn.Module.__init__(self)
    self._build(channels, img_resolution, sr_num_fp16_res, sr_antialias, 3, block_kwargs)


def _sem_ctor(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, semantic_channels, num_fp16_res=4,
              conv_clamp=None, channel_base=None, channel_max=None, **block_kwargs):
    torch.nn.Module.__init__(self)
    self._build(channels, img_resolution, sr_num_fp16_res, sr_antialias, semantic_channels, block_kwargs)


@persistence.persistent_class
class SuperresolutionHybrid8X(_TwoBlockSR):
    """128 -> 512, 128/64 channels (reference :29-57)."""
    OUT_RES, IN_RES, C0, C1, HAS_FILTER_BUFFER = 512, 128, 128, 64, True
    __init__ = _rgb_ctor


@persistence.persistent_class
class SuperresolutionHybrid4X(_TwoBlockSR):
    """128 -> 256 (reference :62-89)."""
    OUT_RES, IN_RES, C0, C1, BLOCK0_UP, RESIZE_ONLY_IF_SMALLER, HAS_FILTER_BUFFER = 256, 128, 128, 64, False, True, True
    __init__ = _rgb_ctor


@persistence.persistent_class
class SuperresolutionHybrid2X(_TwoBlockSR):
    """64 -> 128 (reference :94-122)."""
    OUT_RES, IN_RES, C0, C1, BLOCK0_UP, HAS_FILTER_BUFFER = 128, 64, 128, 64, False, True
    __init__ = _rgb_ctor


@persistence.persistent_class
class SuperresolutionHybrid2X_semantic(_TwoBlockSR):
    """64 -> 128 for the semantic branch (reference :127-155)."""
    OUT_RES, IN_RES, C0, C1, BLOCK0_UP, HAS_FILTER_BUFFER = 128, 64, 128, 64, False, True
    __init__ = _sem_ctor


@persistence.persistent_class
class SuperresolutionHybrid8XDC(_TwoBlockSR):
    """128 -> 512 with 256/128 channels (reference :297-323)."""
    OUT_RES, IN_RES, C0, C1 = 512, 128, 256, 128
    __init__ = _rgb_ctor


@persistence.persistent_class
class SuperresolutionHybrid8XDC_semantic(_TwoBlockSR):
    """128 -> 512 semantic branch with 256/128 channels (reference :328-354)."""
    OUT_RES, IN_RES, C0, C1 = 512, 128, 256, 128
    __init__ = _sem_ctor


@persistence.persistent_class
class SuperresolutionHybridDeepfp32(_TwoBlockSR):
    """Legacy 128 -> 256 variant without antialiasing (reference :160-186)."""
    OUT_RES, IN_RES, C0, C1, BLOCK0_UP, RESIZE_ONLY_IF_SMALLER, HAS_FILTER_BUFFER = 256, 128, 128, 64, False, True, True

    def __init__(self, channels, img_resolution, sr_num_fp16_res, num_fp16_res=4, conv_clamp=None, channel_base=None,
                 channel_max=None, **block_kwargs):
        torch.nn.Module.__init__(self)
        self._build(channels, img_resolution, sr_num_fp16_res, False, 3, block_kwargs)
