"""Loading network pickles written by the reference (`legacy.load_network_pkl`, reference legacy.py:24-59).

The pickles name `torch_utils.persistence._reconstruct_persistent_obj` as the constructor of every persistent object;
with `pix2pix3d_b200.install()` that resolves to this package, which rebuilds the networks as mirror classes (see
torch_utils/persistence.py). TensorFlow-era StyleGAN pickles (the `dnnlib.tflib.network.Network` conversion of reference
legacy.py:61-320) predate pix2pix3D and are not handled.
"""
import copy
import pickle

import torch

from .torch_utils import misc


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == 'dnnlib.tflib.network' and name == 'Network':
            raise NotImplementedError('TensorFlow StyleGAN pickles are not supported; convert them with the reference tooling')
        root = module.split('.', 1)[0]
        if root in ('training', 'torch_utils', 'dnnlib', 'camera_utils', 'legacy'):
            module = f'{__package__}.{module}'          # works with or without install()
        return super().find_class(module, name)


def load_network_pkl(f, force_fp16=False):
    data = _Unpickler(f).load()
    if 'training_set_kwargs' not in data:
        data['training_set_kwargs'] = None
    if 'augment_pipe' not in data:
        data['augment_pipe'] = None
    for key in ('G', 'D', 'G_ema'):
        assert isinstance(data[key], torch.nn.Module)
    assert isinstance(data['training_set_kwargs'], (dict, type(None)))
    assert isinstance(data['augment_pipe'], (torch.nn.Module, type(None)))
    if force_fp16:
        for key in ('G', 'D', 'G_ema'):
            old = data[key]
            kwargs = copy.deepcopy(old.init_kwargs)
            fp16_kwargs = kwargs.get('synthesis_kwargs', kwargs)
            fp16_kwargs.num_fp16_res = 4
            fp16_kwargs.conv_clamp = 256
            if kwargs != old.init_kwargs:
                new = type(old)(**kwargs).eval().requires_grad_(False)
                misc.copy_params_and_buffers(old, new, require_all=True)
                data[key] = new
    return data
