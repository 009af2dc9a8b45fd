"""Conditional (label-map -> 3-D) tri-plane generators of pix2pix3D.

Mirror of the reference's training/triplane_cond.py. The generator classes keep their names, constructor
arguments, sub-module names and method signatures (`mapping(z, c, batch)`, `synthesis(ws, c, ...)`,
`sample`, `sample_mixed`), so callers such as applications/generate_samples.py:113-114 and training/loss.py run
unchanged. `TriPlaneSemanticEntangleGenerator` (:976) is the class every released model uses.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .. import dnnlib, native
from ..torch_utils import misc
from ..torch_utils import persistence
from .networks_stylegan2 import DiscriminatorBlock, FullyConnectedLayer, SynthesisNetwork, normalize_2nd_moment
from .networks_stylegan2 import Generator as StyleGAN2Backbone
from .triplane import OSGDecoder, _decoder_mlp, _mipnerf_sigmoid, _sr_kwargs, fast_synthesis, query_points, render_to_images
from .volumetric_rendering.ray_sampler import RaySampler
from .volumetric_rendering.renderer import ImportanceRenderer, ImportanceSemanticRenderer

# ----------------------------------------------------------------------------------------------
# Label-map encoder and mapping networks
# ----------------------------------------------------------------------------------------------


@persistence.persistent_class
class EqualConv2d(torch.nn.Module):
    """Plain conv with 1/sqrt(fan_in) runtime scaling (:30-61)."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size) * 1.0)
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride = stride
        self.padding = padding
        self.bias = torch.nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        return F.conv2d(input, self.weight * self.scale, bias=self.bias, stride=self.stride, padding=self.padding)

    def __repr__(self):
        return (f'{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]},'
                f' {self.weight.shape[2]}, stride={self.stride}, padding={self.padding})')


@persistence.persistent_class
class Encoder(torch.nn.Module):
    """Discriminator-style pyramid that turns a label map into 'W' or 'W+' latents (:66-196).
    Only the non-progressive configuration the mapping networks instantiate is supported."""

    def __init__(self, img_resolution, img_channels, bottleneck_factor=2, architecture='resnet', channel_base=1,
                 channel_max=512, num_fp16_res=0, conv_clamp=None, lowres_head=None, block_kwargs={}, model_kwargs={},
                 upsample_type='default', progressive=False, **unused):
        super().__init__()
        assert not progressive and lowres_head is None, 'progressive encoders are not part of the pix2pix3D path'
        self.img_resolution = img_resolution
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.img_channels = img_channels
        self.block_resolutions = [2 ** i for i in range(self.img_resolution_log2, bottleneck_factor, -1)]
        self.architecture = architecture
        self.lowres_head = lowres_head
        self.upsample_type = upsample_type
        self.progressive = progressive
        self.model_kwargs = model_kwargs
        self.output_mode = model_kwargs.get('output_mode', 'styles')
        self.predict_camera = model_kwargs.get('predict_camera', False)
        assert not self.predict_camera

        channel_base = int(channel_base * 32768)
        channels = {res: min(channel_base // res, channel_max) for res in self.block_resolutions + [4]}
        fp16_resolution = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        common = dict(img_channels=self.img_channels, architecture=architecture, conv_clamp=conv_clamp)
        cur_layer_idx = 0
        for res in self.block_resolutions:
            block = DiscriminatorBlock(channels[res] if res < img_resolution else 0, channels[res], channels[res // 2],
                                       resolution=res, first_layer_idx=cur_layer_idx, use_fp16=(res >= fp16_resolution),
                                       **block_kwargs, **common)
            setattr(self, f'b{res}', block)
            cur_layer_idx += block.num_layers

        if self.output_mode not in ['W', 'W+', 'None']:
            raise NotImplementedError
        self.num_ws = self.model_kwargs.get('num_ws', 0)
        self.n_latents = self.num_ws if self.output_mode == 'W+' else (0 if self.output_mode == 'None' else 1)
        self.w_dim = self.model_kwargs.get('w_dim', 512)
        self.add_dim = self.model_kwargs.get('add_dim', 0)
        self.out_dim = self.w_dim * self.n_latents + self.add_dim
        assert self.out_dim > 0, 'output dimenstion has to be larger than 0'
        assert self.block_resolutions[-1] // 2 == 4, 'make sure the last resolution is 4x4'
        self.projector = EqualConv2d(channels[4], self.out_dim, 4, padding=0, bias=False)
        self.register_buffer('alpha', torch.scalar_tensor(-1))

    def set_alpha(self, alpha):
        if alpha is not None:
            self.alpha.fill_(alpha)

    def set_resolution(self, res):
        self.curr_status = res

    def forward(self, inputs, **block_kwargs):
        img = inputs['img'] if isinstance(inputs, dict) else inputs
        out = None
        if img.device.type == 'cuda' and not block_kwargs:
            from .. import engine
            if engine.encoder_supported(self, img):
                out = engine.encoder_forward(self, img)
        if out is None:
            x = None
            for res in self.block_resolutions:
                x, img = getattr(self, f'b{res}')(x, img, **block_kwargs)
            out = self.projector(x)[:, :, 0, 0]
        if self.output_mode == 'W+':
            out = out.reshape(out.shape[0], self.num_ws, self.w_dim)
        elif self.output_mode == 'W':
            out = out.unsqueeze(1).repeat(1, self.num_ws, 1)
        else:
            out = None
        return {'ws': out}


class _CondMappingBase(torch.nn.Module):
    """Shared machinery of the four label-conditioned mapping networks (:202-592).

    entangled  (`MaskMappingNetwork`, `EdgeMappingNetwork`):      w = MLP([z, enc(label), embed(c)]) broadcast to num_ws
    disentangled (`*_disentangle`): ws[:, :7] = enc(label) as W+,  ws[:, 7:] = MLP([z, embed(c)]) broadcast
    """
    DISENTANGLE = False
    ENCODER_ATTR = 'embed_mask'
    IS_EDGE = False

    def _setup(self, z_dim, c_dim, in_resolution, in_channels, w_dim, num_ws, num_layers, embed_features, layer_features,
               activation, lr_multiplier, w_avg_beta, one_hot):
        self.z_dim = z_dim
        self.c_dim = c_dim
        self.in_resolution = in_resolution
        self.in_channels = in_channels
        self.w_dim = w_dim
        self.num_ws = num_ws
        self.num_layers = num_layers
        self.w_avg_beta = w_avg_beta
        if not self.IS_EDGE:
            self.one_hot = one_hot
        if self.DISENTANGLE:
            self.geometry_layer = 7
        if embed_features is None:
            embed_features = w_dim
        if layer_features is None:
            layer_features = w_dim
        n_embed = (0 if self.DISENTANGLE else 1) + (1 if c_dim > 0 else 0)
        widths = [z_dim + embed_features * n_embed] + [layer_features] * (num_layers - 1) + [w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, embed_features)
        if self.DISENTANGLE:
            enc_kwargs = {'num_ws': self.geometry_layer, 'w_dim': w_dim, 'output_mode': 'W+'}
        else:
            enc_kwargs = {'num_ws': 1, 'w_dim': embed_features, 'output_mode': 'W'}
        setattr(self, self.ENCODER_ATTR, Encoder(img_resolution=in_resolution, img_channels=in_channels, model_kwargs=enc_kwargs))
        for idx in range(num_layers):
            setattr(self, f'fc{idx}', FullyConnectedLayer(widths[idx], widths[idx + 1], activation=activation,
                                                          lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([num_ws, w_dim] if self.DISENTANGLE else [w_dim]))

    def _label_image(self, batch):
        if self.IS_EDGE:
            return batch['mask'].to(torch.float32)
        if self.one_hot:
            return F.one_hot(batch['mask'].squeeze(1).long(), self.in_channels).permute(0, 3, 1, 2).to(torch.float32)
        return batch['mask'].to(torch.float32)

    def forward(self, z=None, c=None, batch=None, truncation_psi=1, truncation_cutoff=None, update_emas=False, **unused_kwargs):
        encoder = getattr(self, self.ENCODER_ATTR)
        x = None
        if self.z_dim > 0:
            misc.assert_shape(z, [None, self.z_dim])
            x = normalize_2nd_moment(z.to(torch.float32))
        if not self.DISENTANGLE:
            label = self._label_image(batch)
            misc.assert_shape(label, [None, self.in_channels, self.in_resolution, self.in_resolution])
            y = normalize_2nd_moment(encoder(label)['ws'].squeeze(1))
            misc.assert_shape(y, [None, self.w_dim])
            x = torch.cat([x.contiguous(), y.contiguous()], dim=1) if x is not None else y
        if self.c_dim > 0:
            misc.assert_shape(c, [None, self.c_dim])
            c_embed = normalize_2nd_moment(self.embed(c.to(torch.float32)))
            x = torch.cat([x, c_embed], dim=1) if x is not None else c_embed
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)

        if self.DISENTANGLE:
            misc.assert_shape(batch['mask'], [z.shape[0], 1, None, None])
            label = self._label_image(batch)
            misc.assert_shape(label, [z.shape[0], self.in_channels, self.in_resolution, self.in_resolution])
            y = encoder(label)['ws']
            misc.assert_shape(y, [None, self.geometry_layer, self.w_dim])
            if self.num_ws is not None:
                x = torch.cat([y, x.unsqueeze(1).repeat([1, self.num_ws - self.geometry_layer, 1])], dim=1)
            if self.w_avg_beta is not None and update_emas:
                self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        else:
            if self.w_avg_beta is not None and update_emas:
                self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
            if self.num_ws is not None:
                x = x.unsqueeze(1).repeat([1, self.num_ws, 1])

        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


def _mask_ctor(self, z_dim, c_dim, in_resolution, in_channels, w_dim, num_ws, num_layers=8, embed_features=None,
               layer_features=None, activation='lrelu', lr_multiplier=0.01, w_avg_beta=0.995, one_hot=True, **unused):
    torch.nn.Module.__init__(self)
    self._setup(z_dim, c_dim, in_resolution, in_channels, w_dim, num_ws, num_layers, embed_features, layer_features,
                activation, lr_multiplier, w_avg_beta, one_hot)


def _edge_ctor(self, z_dim, c_dim, in_resolution, in_channels, w_dim, num_ws, num_layers=8, embed_features=None,
               layer_features=None, activation='lrelu', lr_multiplier=0.01, w_avg_beta=0.995, **unused):
    torch.nn.Module.__init__(self)
    self._setup(z_dim, c_dim, in_resolution, in_channels, w_dim, num_ws, num_layers, embed_features, layer_features,
                activation, lr_multiplier, w_avg_beta, None)


@persistence.persistent_class
class MaskMappingNetwork(_CondMappingBase):
    """(:202-297)"""
    __init__ = _mask_ctor


@persistence.persistent_class
class MaskMappingNetwork_disentangle(_CondMappingBase):
    """(:301-399) -- the seg2cat / seg2face mapping."""
    DISENTANGLE = True
    __init__ = _mask_ctor


@persistence.persistent_class
class EdgeMappingNetwork(_CondMappingBase):
    """(:404-495)"""
    IS_EDGE, ENCODER_ATTR = True, 'embed_edge'
    __init__ = _edge_ctor


@persistence.persistent_class
class EdgeMappingNetwork_disentangle(_CondMappingBase):
    """(:499-592) -- the edge2car mapping."""
    DISENTANGLE, IS_EDGE = True, True
    __init__ = _edge_ctor


@persistence.persistent_class
class Generator_cond(torch.nn.Module):
    """StyleGAN2 synthesis network + a by-name constructed conditional mapping network (:597-622)."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.z_dim = z_dim
        self.c_dim = c_dim
        self.w_dim = w_dim
        self.img_resolution = img_resolution
        self.img_channels = img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels,
                                          **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = dnnlib.util.construct_class_by_name(**mapping_kwargs, z_dim=z_dim, c_dim=c_dim, w_dim=w_dim,
                                                           num_ws=self.num_ws)

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)


# ----------------------------------------------------------------------------------------------
# Decoders
# ----------------------------------------------------------------------------------------------


class OSGDecoder_semantic(torch.nn.Module):
    """One net; colour outputs optionally passed through the MipNeRF sigmoid (:859-887)."""

    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.net = _decoder_mlp(n_features, self.hidden_dim, 1 + options['decoder_output_dim'], options['decoder_lr_mul'])
        self.final_sigmoid = options['sigmoid']

    def forward(self, sampled_features, ray_directions):
        if native.decoder_mlp_supported(self, sampled_features):      # CUDA: fused forward / backward kernels
            return native.decoder_mlp(self, sampled_features)
        x = sampled_features.mean(1)
        n, m, c = x.shape
        x = self.net(x.view(n * m, c)).view(n, m, -1)
        rgb = _mipnerf_sigmoid(x[..., 1:]) if self.final_sigmoid else x[..., 1:]
        return {'rgb': rgb, 'sigma': x[..., 0:1]}


class OSGDecoder_semantic_entangle(torch.nn.Module):
    """One net whose outputs are [sigma | rgb(3) | semantic logits(Cs) | features] (:891-924)."""

    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.net = _decoder_mlp(n_features, self.hidden_dim, 1 + options['decoder_output_dim'], options['decoder_lr_mul'])
        self.feature_sigmoid = options['sigmoid']
        self.semantic_channels = options['semantic_channels']

    def forward(self, sampled_features, ray_directions):
        if native.decoder_mlp_supported(self, sampled_features):      # CUDA: fused forward / backward kernels
            return native.decoder_mlp(self, sampled_features)
        x = sampled_features.mean(1)
        n, m, c = x.shape
        x = self.net(x.view(n * m, c)).view(n, m, -1)
        if self.feature_sigmoid:
            feature = _mipnerf_sigmoid(x[..., 1:])
        else:
            cs = self.semantic_channels
            feature = torch.cat((_mipnerf_sigmoid(x[..., 1:4]), x[..., 4:4 + cs], _mipnerf_sigmoid(x[..., 4 + cs:])), dim=-1)
        return {'rgb': feature, 'sigma': x[..., 0:1]}


class OSGDecoder_semantic_lateSeparate(torch.nn.Module):
    """Two nets on the same features: colour net and semantic net; density comes from the semantic net (:926-970)."""

    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.net = _decoder_mlp(n_features, self.hidden_dim, 1 + options['decoder_output_dim'], options['decoder_lr_mul'])
        self.net_semantic = _decoder_mlp(n_features, self.hidden_dim, 1 + options['decoder_output_dim'], options['decoder_lr_mul'])
        self.semantic_sigmoid = options['sigmoid']

    def forward(self, sampled_features, ray_directions):
        if native.decoder_mlp_supported(self, sampled_features):      # CUDA: fused forward / backward kernels
            return native.decoder_mlp(self, sampled_features)
        x = sampled_features.mean(1)
        n, m, c = x.shape
        x = x.view(n * m, c)
        rgb = self.net(x).view(n, m, -1)
        semantic = self.net_semantic(x).view(n, m, -1)
        sigma = semantic[..., 0:1]
        rgb = _mipnerf_sigmoid(rgb[..., 1:])
        semantic = _mipnerf_sigmoid(semantic[..., 1:]) if self.semantic_sigmoid else semantic[..., 1:]
        return {'rgb': torch.cat((rgb, semantic), dim=-1), 'sigma': sigma}


# ----------------------------------------------------------------------------------------------
# Generators
# ----------------------------------------------------------------------------------------------


class _CondGeneratorBase(torch.nn.Module):
    """mapping / sample / sample_mixed / forward shared by the conditional generators (:650-660, 1015-1080)."""

    def mapping(self, z, c, batch, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        if self.rendering_kwargs['c_gen_conditioning_zero']:
            c = torch.zeros_like(c)
        return self.backbone.mapping(z, c * self.rendering_kwargs.get('c_scale', 0), batch, truncation_psi=truncation_psi,
                                     truncation_cutoff=truncation_cutoff, update_emas=update_emas)

    def sample(self, coordinates, directions, z, c, batch, truncation_psi=1, truncation_cutoff=None, update_emas=False,
               **synthesis_kwargs):
        ws = self.mapping(z, batch['pose'], batch, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff,
                          update_emas=update_emas)
        return query_points(self, coordinates, directions, ws, update_emas, synthesis_kwargs)

    def sample_mixed(self, coordinates, directions, ws, truncation_psi=1, truncation_cutoff=None, update_emas=False,
                     **synthesis_kwargs):
        return query_points(self, coordinates, directions, ws, update_emas, synthesis_kwargs)

    def forward(self, z, c, batch, truncation_psi=1, truncation_cutoff=None, neural_rendering_resolution=None,
                update_emas=False, cache_backbone=False, use_cached_backbone=False, **synthesis_kwargs):
        ws = self.mapping(z, batch['pose'], batch, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff,
                          update_emas=update_emas)
        return self.synthesis(ws, c, update_emas=update_emas, neural_rendering_resolution=neural_rendering_resolution,
                              cache_backbone=cache_backbone, use_cached_backbone=use_cached_backbone, **synthesis_kwargs)

    def _common_init(self, z_dim, c_dim, w_dim, img_resolution, img_channels, rendering_kwargs):
        self.z_dim = z_dim
        self.c_dim = c_dim
        self.w_dim = w_dim
        self.img_resolution = img_resolution
        self.img_channels = img_channels

    def _sr(self, name, rendering_kwargs, img_resolution, sr_num_fp16_res, sr_kwargs, **extra):
        return dnnlib.util.construct_class_by_name(class_name=rendering_kwargs[name], channels=32, img_resolution=img_resolution,
                                                   sr_num_fp16_res=sr_num_fp16_res, sr_antialias=rendering_kwargs['sr_antialias'],
                                                   **extra, **sr_kwargs)


@persistence.persistent_class
class TriPlaneGenerator(_CondGeneratorBase):
    """Label-conditioned generator without a semantic branch (`--render_mask=False`) (:627-718)."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, sr_num_fp16_res=0, mapping_kwargs={},
                 rendering_kwargs={}, sr_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self._common_init(z_dim, c_dim, w_dim, img_resolution, img_channels, rendering_kwargs)
        self.renderer = ImportanceRenderer()
        self.ray_sampler = RaySampler()
        self.backbone = Generator_cond(z_dim, c_dim, w_dim, img_resolution=256, img_channels=32 * 3,
                                       mapping_kwargs=mapping_kwargs, **synthesis_kwargs)
        self.superresolution = self._sr('superresolution_module', rendering_kwargs, img_resolution, sr_num_fp16_res, sr_kwargs)
        self.decoder = OSGDecoder(32, {'decoder_lr_mul': rendering_kwargs.get('decoder_lr_mul', 1), 'decoder_output_dim': 32})
        self.neural_rendering_resolution = 64
        self.rendering_kwargs = rendering_kwargs
        self._last_planes = None

    def synthesis(self, ws, c, neural_rendering_resolution=None, update_emas=False, cache_backbone=False,
                  use_cached_backbone=False, **synthesis_kwargs):
        fast = fast_synthesis(self, ws, c, neural_rendering_resolution, cache_backbone, use_cached_backbone, synthesis_kwargs)
        if fast is not None:
            return fast
        feature_image, depth_image = render_to_images(self, ws, c, neural_rendering_resolution, update_emas,
                                                      cache_backbone, use_cached_backbone, synthesis_kwargs)
        rgb_image = feature_image[:, :3]
        sr_image = self.superresolution(rgb_image, feature_image, ws,
                                        noise_mode=self.rendering_kwargs['superresolution_noise_mode'],
                                        **_sr_kwargs(synthesis_kwargs))
        return {'image': sr_image, 'image_raw': rgb_image, 'image_depth': depth_image}


def _semantic_heads(gen, feature_image, ws, synthesis_kwargs):
    """Split the 64-channel feature image into colour / semantic halves and super-resolve both (:1052-1059)."""
    half = feature_image.shape[1] // 2
    rgb_feat, sem_feat = feature_image[:, :half], feature_image[:, half:]
    noise_mode = gen.rendering_kwargs['superresolution_noise_mode']
    rgb_image = rgb_feat[:, :3]
    sr_image = gen.superresolution(rgb_image, rgb_feat, ws, noise_mode=noise_mode, **_sr_kwargs(synthesis_kwargs))
    sem_image = sem_feat[:, :gen.semantic_channels]
    sr_sem = gen.superresolution_semantic(sem_image, sem_feat, ws, noise_mode=noise_mode, **_sr_kwargs(synthesis_kwargs))
    return rgb_image, sr_image, sem_image, sr_sem


@persistence.persistent_class
class TriPlaneSemanticGenerator(torch.nn.Module):
    """Separate texture and semantic tri-plane backbones rendered together by ImportanceSemanticRenderer (:724-849).
    `ws` carries both latents side by side: [..., :w_dim] texture, [..., w_dim:] semantic."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, semantic_channels, sr_num_fp16_res=0,
                 mapping_kwargs={}, rendering_kwargs={}, sr_kwargs={}, data_type=None, **synthesis_kwargs):
        super().__init__()
        self.z_dim = z_dim
        self.c_dim = c_dim
        self.w_dim = w_dim
        self.img_resolution = img_resolution
        self.img_channels = img_channels
        self.semantic_channels = semantic_channels
        self.data_type = data_type
        self.renderer = ImportanceSemanticRenderer()
        self.ray_sampler = RaySampler()
        self.backbone = StyleGAN2Backbone(z_dim, c_dim, w_dim, img_resolution=256, img_channels=32 * 3,
                                          mapping_kwargs=mapping_kwargs, **synthesis_kwargs)
        self.backbone_semantic = Generator_cond(0, c_dim, w_dim, img_resolution=256, img_channels=32 * 3,
                                                mapping_kwargs=mapping_kwargs, **synthesis_kwargs)
        self.superresolution = dnnlib.util.construct_class_by_name(
            class_name=rendering_kwargs['superresolution_module'], channels=32, img_resolution=img_resolution,
            sr_num_fp16_res=sr_num_fp16_res, sr_antialias=rendering_kwargs['sr_antialias'], **sr_kwargs)
        self.superresolution_semantic = dnnlib.util.construct_class_by_name(
            class_name=rendering_kwargs['superresolution_module_semantic'], channels=32, img_resolution=img_resolution,
            sr_num_fp16_res=sr_num_fp16_res, sr_antialias=rendering_kwargs['sr_antialias'],
            semantic_channels=semantic_channels, **sr_kwargs)
        lr_mul = rendering_kwargs.get('decoder_lr_mul', 1)
        self.decoder = OSGDecoder(64, {'decoder_lr_mul': lr_mul, 'decoder_output_dim': 32, 'sigmoid': True})
        self.decoder_semantic = OSGDecoder_semantic(32, {'decoder_lr_mul': lr_mul, 'decoder_output_dim': 32,
                                                         'sigmoid': True if semantic_channels == 1 else False})
        self.neural_rendering_resolution = 64
        self.rendering_kwargs = rendering_kwargs
        self._last_planes = None

    def mapping(self, z, c, batch, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        if self.rendering_kwargs['c_gen_conditioning_zero']:
            c = torch.zeros_like(c)
        cs = c * self.rendering_kwargs.get('c_scale', 0)
        ws_texture = self.backbone.mapping(z, cs, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff,
                                           update_emas=update_emas)
        ws_semantic = self.backbone_semantic.mapping(None, cs, batch, truncation_psi=truncation_psi,
                                                     truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return torch.cat([ws_texture, ws_semantic], dim=-1)

    def _planes(self, ws, update_emas, synthesis_kwargs):
        assert ws.shape[-1] == self.w_dim * 2
        ws_texture, ws_semantic = ws[..., :self.w_dim], ws[..., self.w_dim:]
        pt = self.backbone.synthesis(ws_texture, update_emas=update_emas, **synthesis_kwargs)
        ps = self.backbone_semantic.synthesis(ws_semantic, update_emas=update_emas, **synthesis_kwargs)
        pt = pt.view(len(pt), 3, 32, pt.shape[-2], pt.shape[-1])
        ps = ps.view(len(ps), 3, 32, ps.shape[-2], ps.shape[-1])
        return ws_texture, ws_semantic, pt, ps

    def synthesis(self, ws, c, neural_rendering_resolution=None, update_emas=False, cache_backbone=False,
                  use_cached_backbone=False, **synthesis_kwargs):
        cam2world = c[:, :16].view(-1, 4, 4)
        intrinsics = c[:, 16:25].view(-1, 3, 3)
        if neural_rendering_resolution is None:
            neural_rendering_resolution = self.neural_rendering_resolution
        else:
            self.neural_rendering_resolution = neural_rendering_resolution
        ray_origins, ray_directions = self.ray_sampler(cam2world, intrinsics, neural_rendering_resolution)
        n = ray_origins.shape[0]
        ws_texture, ws_semantic, planes_texture, planes_semantic = self._planes(ws, update_emas, synthesis_kwargs)
        feats, depth, _ = self.renderer(planes_texture, planes_semantic, self.decoder, self.decoder_semantic, ray_origins,
                                        ray_directions, self.rendering_kwargs)
        h = w = self.neural_rendering_resolution
        feature_image = feats.permute(0, 2, 1).reshape(n, feats.shape[-1], h, w).contiguous()
        depth_image = depth.permute(0, 2, 1).reshape(n, 1, h, w)
        half = feature_image.shape[1] // 2
        rgb_feat, sem_feat = feature_image[:, :half], feature_image[:, half:]
        noise_mode = self.rendering_kwargs['superresolution_noise_mode']
        rgb_image = rgb_feat[:, :3]
        sr_image = self.superresolution(rgb_image, rgb_feat, ws_texture, noise_mode=noise_mode, **_sr_kwargs(synthesis_kwargs))
        sem_image = sem_feat[:, :self.semantic_channels]
        sr_sem = self.superresolution_semantic(sem_image, sem_feat, ws_semantic, noise_mode=noise_mode,
                                               **_sr_kwargs(synthesis_kwargs))
        return {'image': sr_image, 'image_raw': rgb_image, 'image_depth': depth_image, 'semantic': sr_sem,
                'semantic_raw': sem_image}

    def sample(self, coordinates, directions, z, c, batch, truncation_psi=1, truncation_cutoff=None, update_emas=False,
               **synthesis_kwargs):
        ws = self.mapping(z, batch['pose'], batch, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff,
                          update_emas=update_emas)
        return self.sample_mixed(coordinates, directions, ws, update_emas=update_emas, **synthesis_kwargs)

    def sample_mixed(self, coordinates, directions, ws, truncation_psi=1, truncation_cutoff=None, update_emas=False,
                     **synthesis_kwargs):
        _, _, planes_texture, planes_semantic = self._planes(ws, update_emas, synthesis_kwargs)
        return self.renderer.run_model(planes_texture, planes_semantic, self.decoder, self.decoder_semantic, coordinates,
                                       directions, self.rendering_kwargs)

    def forward(self, z, c, batch, truncation_psi=1, truncation_cutoff=None, neural_rendering_resolution=None,
                update_emas=False, cache_backbone=False, use_cached_backbone=False, **synthesis_kwargs):
        ws = self.mapping(z, batch['pose'], batch, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff,
                          update_emas=update_emas)
        return self.synthesis(ws, c, update_emas=update_emas, neural_rendering_resolution=neural_rendering_resolution,
                              cache_backbone=cache_backbone, use_cached_backbone=use_cached_backbone, **synthesis_kwargs)


@persistence.persistent_class
class TriPlaneSemanticEntangleGenerator(_CondGeneratorBase):
    """Colour + semantics from one tri-plane set via the two-headed decoder (:976-1080)."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, semantic_channels, sr_num_fp16_res=0,
                 mapping_kwargs={}, rendering_kwargs={}, sr_kwargs={}, data_type=None, **synthesis_kwargs):
        super().__init__()
        self._common_init(z_dim, c_dim, w_dim, img_resolution, img_channels, rendering_kwargs)
        self.semantic_channels = semantic_channels
        self.data_type = data_type
        self.renderer = ImportanceRenderer()
        self.ray_sampler = RaySampler()
        self.backbone = Generator_cond(z_dim, c_dim, w_dim, img_resolution=256, img_channels=32 * 3,
                                       mapping_kwargs=mapping_kwargs, **synthesis_kwargs)
        self.superresolution = self._sr('superresolution_module', rendering_kwargs, img_resolution, sr_num_fp16_res, sr_kwargs)
        self.superresolution_semantic = self._sr('superresolution_module_semantic', rendering_kwargs, img_resolution,
                                                 sr_num_fp16_res, sr_kwargs, semantic_channels=semantic_channels)
        self.decoder = OSGDecoder_semantic_lateSeparate(32, {
            'decoder_lr_mul': rendering_kwargs.get('decoder_lr_mul', 1), 'decoder_output_dim': 32,
            'sigmoid': True if semantic_channels == 1 else False, 'semantic_channels': semantic_channels})
        self.neural_rendering_resolution = 64
        self.rendering_kwargs = rendering_kwargs
        self._last_planes = None

    def synthesis(self, ws, c, neural_rendering_resolution=None, update_emas=False, cache_backbone=False,
                  use_cached_backbone=False, **synthesis_kwargs):
        fast = fast_synthesis(self, ws, c, neural_rendering_resolution, cache_backbone, use_cached_backbone, synthesis_kwargs)
        if fast is not None:
            return fast
        feature_image, depth_image = render_to_images(self, ws, c, neural_rendering_resolution, update_emas,
                                                      cache_backbone, use_cached_backbone, synthesis_kwargs)
        rgb_image, sr_image, sem_image, sr_sem = _semantic_heads(self, feature_image, ws, synthesis_kwargs)
        return {'image': sr_image, 'image_raw': rgb_image, 'image_depth': depth_image, 'semantic': sr_sem,
                'semantic_raw': sem_image}


@persistence.persistent_class
class TriPlaneSemanticEntangleGenerator_withBG(_CondGeneratorBase):
    """Entangle generator plus a spherical background plane blended behind the volume (:1085-1246)."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, semantic_channels, sr_num_fp16_res=0,
                 mapping_kwargs={}, rendering_kwargs={}, sr_kwargs={}, data_type=None, **synthesis_kwargs):
        super().__init__()
        self._common_init(z_dim, c_dim, w_dim, img_resolution, img_channels, rendering_kwargs)
        self.semantic_channels = semantic_channels
        self.data_type = data_type
        self.renderer = ImportanceRenderer()
        self.ray_sampler = RaySampler()
        self.backbone = Generator_cond(z_dim, c_dim, w_dim, img_resolution=256, img_channels=32 * 3,
                                       mapping_kwargs=mapping_kwargs, **synthesis_kwargs)
        mapping_bg_kwargs = mapping_kwargs.copy()
        mapping_bg_kwargs['class_name'] = None
        self.backbone_bg = StyleGAN2Backbone(z_dim, 0, w_dim, img_resolution=256, img_channels=32 * 2,
                                             mapping_kwargs=mapping_bg_kwargs, **synthesis_kwargs)
        self.superresolution = self._sr('superresolution_module', rendering_kwargs, img_resolution, sr_num_fp16_res, sr_kwargs)
        self.superresolution_semantic = self._sr('superresolution_module_semantic', rendering_kwargs, img_resolution,
                                                 sr_num_fp16_res, sr_kwargs, semantic_channels=semantic_channels)
        self.decoder = OSGDecoder_semantic_lateSeparate(32, {
            'decoder_lr_mul': rendering_kwargs.get('decoder_lr_mul', 1), 'decoder_output_dim': 32,
            'sigmoid': True if semantic_channels == 1 else False, 'semantic_channels': semantic_channels})
        self.neural_rendering_resolution = 64
        self.rendering_kwargs = rendering_kwargs
        self._last_planes = None

    def synthesis(self, ws, c, neural_rendering_resolution=None, update_emas=False, cache_backbone=False,
                  use_cached_backbone=False, **synthesis_kwargs):
        cam2world = c[:, :16].view(-1, 4, 4)
        intrinsics = c[:, 16:25].view(-1, 3, 3)
        if neural_rendering_resolution is None:
            neural_rendering_resolution = self.neural_rendering_resolution
        else:
            self.neural_rendering_resolution = neural_rendering_resolution
        ray_origins, ray_directions = self.ray_sampler(cam2world, intrinsics, neural_rendering_resolution)
        n = ray_origins.shape[0]
        if use_cached_backbone and self._last_planes is not None:
            planes = self._last_planes
        else:
            planes = self.backbone.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)
        if cache_backbone:
            self._last_planes = planes
        planes = planes.view(len(planes), 3, 32, planes.shape[-2], planes.shape[-1])
        feats, depth, wsum = self.renderer(planes, self.decoder, ray_origins, ray_directions, self.rendering_kwargs)

        ws_bg = ws[:, -1, :].unsqueeze(1).repeat([1, ws.shape[1], 1])
        planes_bg = self.backbone_bg.synthesis(ws_bg, update_emas=update_emas, **synthesis_kwargs)
        planes_bg = planes_bg.view(len(planes_bg), 64, planes_bg.shape[-2], planes_bg.shape[-1])
        feats, depth = self.combine_fg_bg(feats, depth, wsum, planes_bg, ray_origins, ray_directions, self.rendering_kwargs)

        h = w = self.neural_rendering_resolution
        feature_image = feats.permute(0, 2, 1).reshape(n, feats.shape[-1], h, w).contiguous()
        depth_image = depth.permute(0, 2, 1).reshape(n, 1, h, w)
        weight_image = wsum.permute(0, 2, 1).reshape(n, 1, h, w)
        rgb_image, sr_image, sem_image, sr_sem = _semantic_heads(self, feature_image, ws, synthesis_kwargs)
        return {'image': sr_image, 'image_raw': rgb_image, 'image_depth': depth_image, 'semantic': sr_sem,
                'semantic_raw': sem_image, 'weight': weight_image}

    def combine_fg_bg(self, feature_samples, depth_samples, weights_samples, planes_bg, ray_origins, ray_directions,
                      rendering_kwargs):
        """Look the background up by ray direction (spherical coordinates) and composite it behind (:1207-1246)."""
        d = ray_directions / torch.norm(ray_directions, dim=-1, keepdim=True)
        theta = torch.atan2(d[:, :, 1], d[:, :, 0])
        phi = torch.acos(d[:, :, 2])
        grid = torch.stack([theta * 2 / np.pi, phi * 2 / np.pi - 1], dim=-1).unsqueeze(1)
        bg = F.grid_sample(planes_bg, grid, mode='bilinear', padding_mode='border')
        bg = bg.squeeze(2).permute(0, 2, 1)
        assert bg.shape == feature_samples.shape
        bg = _mipnerf_sigmoid(bg) * 2 - 1
        bg[:, :, 32:] = bg[:, :, 32:] * 10
        if self.semantic_channels > 1:
            bg[:, :, 32 + 1:32 + self.semantic_channels] = 0
            bg[:, :, 32] = 20
        feature_samples = feature_samples + bg * (1 - weights_samples)
        depth_bg = torch.ones_like(depth_samples) * rendering_kwargs['ray_end']
        depth_samples = depth_samples + depth_bg * (1 - weights_samples)
        return feature_samples, depth_samples


from .triplane import mark_first_order  # noqa: E402

mark_first_order(TriPlaneGenerator, TriPlaneSemanticGenerator, TriPlaneSemanticEntangleGenerator, TriPlaneSemanticEntangleGenerator_withBG)
